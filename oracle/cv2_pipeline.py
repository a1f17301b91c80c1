"""Second, independent restatement of the reference hot path that calls the REAL OpenCV primitives (cv2 4.13).

TEST INFRASTRUCTURE ONLY.  Purpose: pin oracle/ivslam_oracle.cpp.  The reference's pixel arithmetic is OpenCV's
(cv::resize, cv::FAST, cv::GaussianBlur, cv::fastAtan2; call sites src/ORBextractor.cc:1311, :1045/:1051, :1277, :104),
and OpenCV is not vendored in the reference; this driver follows ORBextractor::operator() (src/ORBextractor.cc:1224-1296),
ComputePyramid (:1298-1323), ComputeKeyPointsOld (:880-1213), IC_Angle (:78-105) and computeOrbDescriptor (:108-148) in
Python, composing cv2 calls exactly where the reference calls OpenCV.  The only non-cv2 pieces are
std::nth_element (inside KeyPointsFilter::retainBest, not exposed by cv2) — taken from the real libstdc++ through
oracle_lib.retain_best — and glibc cosf/sinf through ctypes.

It is slow (Python loops) and is used for golden-vector generation and small/medium parity checks only.
"""
import ctypes
import math

import cv2
import numpy as np

from . import oracle_lib

f32 = np.float32
_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

EDGE = 19
HALF_PATCH = 15
PATCH = 31


def _load_pattern():
    import os
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(here, "..", "include", "ivslam_brief_pattern.inc")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    v = np.array([int(t) for t in re.findall(r"-?\d+", txt)], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


PATTERN = _load_pattern()


def cv_round(x):
    return int(np.rint(x))


class Cv2Extractor:
    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, enableIntrospection=False):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.iniTh, self.minTh, self.intro = iniThFAST, minThFAST, enableIntrospection
        sf = float(f32(scaleFactor))     # member is double, initialised from the float argument
        self.scale = [f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sf))
        self.inv = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sf)
        nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.nper, tot = [], 0
        for _ in range(nlevels - 1):
            self.nper.append(cv_round(nd))
            tot += self.nper[-1]
            nd = f32(nd * factor)
        self.nper.append(max(nfeatures - tot, 0))
        um = [0] * (HALF_PATCH + 1)
        vmax = int(math.floor(f32(HALF_PATCH) * f32(math.sqrt(2.0)) / f32(2) + f32(1)))
        vmin = int(math.ceil(f32(HALF_PATCH) * f32(math.sqrt(2.0)) / f32(2)))
        for v in range(vmax + 1):
            um[v] = cv_round(math.sqrt(HALF_PATCH * HALF_PATCH - v * v))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while um[v0] == um[v0 + 1]:
                v0 += 1
            um[v] = v0
            v0 += 1
        self.umax = um
        self.fast = {th: cv2.FastFeatureDetector_create(th, True) for th in (iniThFAST, minThFAST)}

    def pyramid(self, img):
        h, w = img.shape
        out = []
        for l in range(self.nlevels):
            sz = (cv_round(f32(w) * self.inv[l]), cv_round(f32(h) * self.inv[l]))
            out.append(img.copy() if l == 0 else cv2.resize(out[-1], sz, interpolation=cv2.INTER_LINEAR))
        return out

    def _fast(self, win, th):
        kps = self.fast[th].detect(np.ascontiguousarray(win))
        return [(kp.pt[0], kp.pt[1], kp.response) for kp in kps]

    def keypoints_old(self, pyr, qpyr):
        weighted = qpyr is not None and self.intro
        ratio = f32(pyr[0].shape[1]) / f32(pyr[0].shape[0])
        allk = []
        for level in range(self.nlevels):
            img = pyr[level]
            H_, W_ = img.shape
            nDes = self.nper[level]
            levelCols = int(np.sqrt(f32(nDes) / (f32(5) * ratio), dtype=f32))
            levelRows = int(ratio * f32(levelCols))
            maxBX, maxBY = W_ - EDGE, H_ - EDGE
            W, H = maxBX - EDGE, maxBY - EDGE
            cellW = int(np.ceil(f32(W) / f32(levelCols)))
            cellH = int(np.ceil(f32(H) / f32(levelRows)))
            nCells = levelRows * levelCols
            nfeaturesCell = int(np.ceil(f32(nDes) / f32(nCells)))
            cells = [[None] * levelCols for _ in range(levelRows)]
            nToRetain = np.zeros((levelRows, levelCols), np.int64)
            nTotal = np.zeros((levelRows, levelCols), np.int64)
            noMore = np.zeros((levelRows, levelCols), bool)
            iniXCol, iniYRow = [0] * levelCols, [0] * levelRows
            nNoMore = nToDist = 0
            hY = f32(cellH + 6)
            nf_cell = np.full((levelRows, levelCols), f32(nfeaturesCell), f32)
            cw = np.zeros((levelRows, levelCols), f32)
            cw_sum = f32(0)
            if weighted:
                q = qpyr[level]
                for i in range(levelRows):
                    iniY = f32(EDGE + i * cellH - 3)
                    iniYRow[i] = int(iniY)
                    if i == levelRows - 1:
                        hY = f32(maxBY + 3) - iniY
                        if hY <= 0:
                            continue
                    hX = f32(cellW + 6)
                    for j in range(levelCols):
                        if i == 0:
                            iniX = f32(EDGE + j * cellW - 3)
                            iniXCol[j] = int(iniX)
                        else:
                            iniX = f32(iniXCol[j])
                        if j == levelCols - 1:
                            hX = f32(maxBX + 3) - iniX
                            if hX <= 0:
                                continue
                        roi = q[int(iniY):int(iniY + hY), int(iniX):int(iniX + hX)]
                        s = int(cv2.sumElems(np.ascontiguousarray(roi))[0])
                        cost = f32(s) / f32(hX * hY)
                        qs = f32(1.0 / (1.0 + float(cost / f32(255))))
                        qn = f32(f32(2) * qs - f32(1))
                        cw[i, j] = qn
                        cw_sum = f32(cw_sum + qn)
            for i in range(levelRows):
                iniY = f32(EDGE + i * cellH - 3)
                iniYRow[i] = int(iniY)
                if i == levelRows - 1:
                    hY = f32(maxBY + 3) - iniY
                    if hY <= 0:
                        continue
                hX = f32(cellW + 6)
                for j in range(levelCols):
                    if i == 0:
                        iniX = f32(EDGE + j * cellW - 3)
                        iniXCol[j] = int(iniX)
                    else:
                        iniX = f32(iniXCol[j])
                    if j == levelCols - 1:
                        hX = f32(maxBX + 3) - iniX
                        if hX <= 0:
                            continue
                    if weighted:
                        with np.errstate(invalid="ignore", divide="ignore"):
                            v = np.ceil(f32(f32(nDes) * cw[i, j]) / cw_sum)
                        nf_cell[i, j] = f32(1.0) if not (f32(1.0) < v) else v     # std::max(1.0f, v)
                    y0, y1, x0, x1 = int(iniY), int(iniY + hY), int(iniX), int(iniX + hX)
                    win = img[y0:y1, x0:x1]
                    k = self._fast(win, self.iniTh)
                    if len(k) <= 3:
                        k = self._fast(win, self.minTh)
                    resp = np.array([t[2] for t in k], f32)
                    if weighted:
                        q = qpyr[level]
                        for n_, t in enumerate(k):
                            cost = f32(q[y0 + int(t[1]), x0 + int(t[0])])
                            resp[n_] = f32(resp[n_] * f32(f32(2) * (f32(1.0) / (f32(1.0) + cost / f32(255.0))) - f32(1)))
                    cells[i][j] = (k, resp)
                    nKeys = len(k)
                    nTotal[i, j] = nKeys
                    if f32(nKeys) > nf_cell[i, j]:
                        nToRetain[i, j] = int(nf_cell[i, j])
                        noMore[i, j] = False
                    else:
                        nToRetain[i, j] = nKeys
                        nToDist = int(f32(nToDist) + (nf_cell[i, j] - f32(nKeys)))
                        noMore[i, j] = True
                        nNoMore += 1
            while nToDist > 0 and nNoMore < nCells:
                for i in range(levelRows):
                    for j in range(levelCols):
                        if not noMore[i, j]:
                            nNew = int(nf_cell[i, j] + np.ceil(f32(nToDist) / f32(nCells - nNoMore)))
                            if nTotal[i, j] > nNew:
                                nToRetain[i, j] = nNew
                                noMore[i, j] = False
                            else:
                                nToRetain[i, j] = nTotal[i, j]
                                nToDist += nNew - int(nTotal[i, j])
                                noMore[i, j] = True
                                nNoMore += 1
                nToDist = 0
            size = f32(int(f32(PATCH) * self.scale[level]))
            lx, ly, lr = [], [], []
            for i in range(levelRows):
                for j in range(levelCols):
                    if cells[i][j] is None:
                        continue
                    k, resp = cells[i][j]
                    keep = oracle_lib.retain_best(resp, int(nToRetain[i, j])) if len(k) else []
                    for idx in keep:
                        lx.append(f32(k[idx][0]) + f32(iniXCol[j]))
                        ly.append(f32(k[idx][1]) + f32(iniYRow[i]))
                        lr.append(resp[idx])
            lx, ly, lr = np.array(lx, f32), np.array(ly, f32), np.array(lr, f32)
            if lx.size > nDes:
                keep = oracle_lib.retain_best(lr, nDes)
                lx, ly, lr = lx[keep], ly[keep], lr[keep]
            allk.append(dict(x=lx, y=ly, response=lr, size=size))
        for level in range(self.nlevels):
            k = allk[level]
            k["angle"] = np.array([self.ic_angle(pyr[level], k["x"][i], k["y"][i]) for i in range(k["x"].size)], f32)
        return allk

    # ---- optional mode: ComputeKeyPointsOctTree + DistributeOctTree + DivideNode (src/ORBextractor.cc:771-878, :545-769,
    # :487-543; dead code in the reference).  Ties in the (size, node) sort are broken by creation order (the reference
    # uses heap addresses) — the same deterministic rule as oracle/ivslam_oracle.cpp.
    def keypoints_octree(self, pyr):
        allk = []
        for level in range(self.nlevels):
            img = pyr[level]
            H_, W_ = img.shape
            minB = EDGE - 3
            maxBX, maxBY = W_ - EDGE + 3, H_ - EDGE + 3
            width, height = f32(maxBX - minB), f32(maxBY - minB)
            nCols, nRows = int(width / f32(30)), int(height / f32(30))
            wCell, hCell = int(np.ceil(width / f32(nCols))), int(np.ceil(height / f32(nRows)))
            keys = []      # (x, y, response) relative to (minB, minB)
            for i in range(nRows):
                iniY = minB + i * hCell
                maxY = iniY + hCell + 6
                if iniY >= maxBY - 3:
                    continue
                maxY = min(maxY, maxBY)
                for j in range(nCols):
                    iniX = minB + j * wCell
                    maxX = iniX + wCell + 6
                    if iniX >= maxBX - 6:
                        continue
                    maxX = min(maxX, maxBX)
                    win = img[iniY:maxY, iniX:maxX]
                    k = self._fast(win, self.iniTh)
                    if not k:
                        k = self._fast(win, self.minTh)
                    keys += [(f32(t[0]) + f32(j * wCell), f32(t[1]) + f32(i * hCell), f32(t[2])) for t in k]
            sel = self._distribute_octree(keys, minB, maxBX, minB, maxBY, self.nper[level])
            lx = np.array([keys[i][0] + f32(minB) for i in sel], f32)
            ly = np.array([keys[i][1] + f32(minB) for i in sel], f32)
            lr = np.array([keys[i][2] for i in sel], f32)
            allk.append(dict(x=lx, y=ly, response=lr, size=f32(int(f32(PATCH) * self.scale[level]))))
        for level in range(self.nlevels):
            k = allk[level]
            k["angle"] = np.array([self.ic_angle(pyr[level], k["x"][i], k["y"][i]) for i in range(k["x"].size)], f32)
        return allk

    @staticmethod
    def _distribute_octree(K, minX, maxX, minY, maxY, N):
        nIni = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))
        hX = f32(maxX - minX) / f32(nIni)
        seq = [0]

        def node(ulx, uly, brx, bry, keys):
            seq[0] += 1
            return dict(ulx=ulx, uly=uly, brx=brx, bry=bry, keys=keys, noMore=len(keys) == 1, seq=seq[0])

        def divide(n):
            halfX = int(math.ceil(float(f32(n["brx"] - n["ulx"]) / f32(2))))
            halfY = int(math.ceil(float(f32(n["bry"] - n["uly"]) / f32(2))))
            mx, my = n["ulx"] + halfX, n["uly"] + halfY
            ks = [[], [], [], []]
            for k in n["keys"]:
                x, y = K[k][0], K[k][1]
                if x < mx:
                    ks[0 if y < my else 2].append(k)
                else:
                    ks[1 if y < my else 3].append(k)
            boxes = [(n["ulx"], n["uly"], mx, my), (mx, n["uly"], n["brx"], my), (n["ulx"], my, mx, n["bry"]), (mx, my, n["brx"], n["bry"])]
            return [node(*boxes[c], ks[c]) for c in range(4)]

        ini = [node(int(hX * f32(i)), 0, int(hX * f32(i + 1)), maxY - minY, []) for i in range(nIni)]
        for i, kp in enumerate(K):
            ini[int(kp[0] / hX)]["keys"].append(i)
        nodes = []          # python list used as std::list: index 0 = front
        for n in ini:
            if len(n["keys"]) == 1:
                n["noMore"] = True
                nodes.append(n)
            elif n["keys"]:
                n["noMore"] = False
                nodes.append(n)
        finish = False
        while not finish:
            prevSize = len(nodes)
            nToExpand = 0
            expand = []
            i = 0
            while i < len(nodes):
                n = nodes[i]
                if n["noMore"]:
                    i += 1
                    continue
                for c in divide(n):
                    if c["keys"]:
                        nodes.insert(0, c)
                        i += 1
                        if len(c["keys"]) > 1:
                            nToExpand += 1
                            expand.append(c)
                del nodes[i]
            if len(nodes) >= N or len(nodes) == prevSize:
                finish = True
            elif len(nodes) + nToExpand * 3 > N:
                while not finish:
                    prev2 = len(nodes)
                    prev = sorted(expand, key=lambda n: (len(n["keys"]), n["seq"]))
                    expand = []
                    for n in reversed(prev):
                        for c in divide(n):
                            if c["keys"]:
                                nodes.insert(0, c)
                                if len(c["keys"]) > 1:
                                    expand.append(c)
                        nodes.remove(n)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev2:
                        finish = True
        out = []
        for n in nodes:
            best = n["keys"][0]
            for k in n["keys"][1:]:
                if K[k][2] > K[best][2]:
                    best = k
            out.append(best)
        return out

    def ic_angle(self, img, px, py):
        cx, cy = cv_round(px), cv_round(py)
        m01 = m10 = 0
        for v in range(-HALF_PATCH, HALF_PATCH + 1):
            d = self.umax[abs(v)]
            row = img[cy + v, cx - d:cx + d + 1].astype(np.int64)
            m10 += int((np.arange(-d, d + 1) * row).sum())
            m01 += v * int(row.sum())
        return f32(cv2.fastAtan2(float(f32(m01)), float(f32(m10))))

    def descriptor(self, blur, px, py, angle_deg):
        factorPI = f32(math.pi / 180.0)
        ang = f32(f32(angle_deg) * factorPI)
        a, b = f32(_libm.cosf(float(ang))), f32(_libm.sinf(float(ang)))
        cx, cy = cv_round(px), cv_round(py)
        X, Y = PATTERN[:, 0].astype(f32), PATTERN[:, 1].astype(f32)
        ry = np.rint(X * b + Y * a).astype(np.int64)       # float32 mul, float32 add, round-half-even
        rx = np.rint(X * a - Y * b).astype(np.int64)
        vals = blur[cy + ry, cx + rx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits, bitorder="little")

    def __call__(self, image, mask=None):
        qpyr = self.pyramid(mask) if (mask is not None and self.intro) else None
        pyr = self.pyramid(image)
        allk = self.keypoints_octree(pyr) if getattr(self, "kp_mode", 0) == 1 else self.keypoints_old(pyr, qpyr)
        kps, descs, blurs = [], [], []
        for level in range(self.nlevels):
            k = allk[level]
            n = k["x"].size
            if n == 0:
                blurs.append(None)
                continue
            blur = cv2.GaussianBlur(pyr[level].copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            blurs.append(blur)
            rec = np.zeros(n, oracle_lib.KP_DTYPE)
            for i in range(n):
                descs.append(self.descriptor(blur, k["x"][i], k["y"][i], k["angle"][i]))
            s = self.scale[level]
            rec["x"] = k["x"] * s if level else k["x"]
            rec["y"] = k["y"] * s if level else k["y"]
            rec["size"], rec["angle"], rec["response"], rec["octave"], rec["class_id"] = k["size"], k["angle"], k["response"], level, -1
            kps.append(rec)
        self.last_pyramid, self.last_blur, self.last_qpyr = pyr, blurs, qpyr
        if not kps:
            return np.zeros(0, oracle_lib.KP_DTYPE), np.zeros((0, 32), np.uint8)
        return np.concatenate(kps), np.stack(descs)


def hamming(a, b):
    """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1700-1716)."""
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _c_round(x):
    """C round(): half away from zero, on a float32 value (src/Frame.cc:849-851)."""
    x = float(x)
    return f32(math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5))


def compute_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale, mbf, maxD):
    """Frame::ComputeStereoMatches (src/Frame.cc:758-932) with cv2.norm for the SAD, maxD explicit."""
    N = kL.size
    uRight = np.full(N, -1, f32)
    depth = np.full(N, -1, f32)
    mbf, maxD = f32(mbf), f32(maxD)
    thOrbDist = (100 + 50) // 2
    nRows = pyrL[0].shape[0]
    rows = [[] for _ in range(nRows)]
    for iR in range(kR.size):
        kpY = f32(kR["y"][iR])
        r = f32(2.0) * scale[kR["octave"][iR]]
        maxr, minr = int(np.ceil(kpY + r)), int(np.floor(kpY - r))
        for yi in range(minr, maxr + 1):
            if 0 <= yi < nRows:
                rows[yi].append(iR)
    minD = f32(0)
    vDistIdx = []
    for iL in range(N):
        levelL = int(kL["octave"][iL])
        vL, uL = f32(kL["y"][iL]), f32(kL["x"][iL])
        cand = rows[int(vL)]
        if not cand:
            continue
        minU, maxU = uL - maxD, uL - minD
        if maxU < 0:
            continue
        bestDist, bestIdxR = 100, 0
        for iR in cand:
            o = int(kR["octave"][iR])
            if o < levelL - 1 or o > levelL + 1:
                continue
            uR = f32(kR["x"][iR])
            if minU <= uR <= maxU:
                d = hamming(dL[iL], dR[iR])
                if d < bestDist:
                    bestDist, bestIdxR = d, iR
        if bestDist < thOrbDist:
            uR0 = f32(kR["x"][bestIdxR])
            sf = inv_scale[levelL]
            suL, svL, suR0 = _c_round(uL * sf), _c_round(vL * sf), _c_round(uR0 * sf)
            w = L = 5
            IL = pyrL[levelL][int(svL) - w:int(svL) + w + 1, int(suL) - w:int(suL) + w + 1].astype(f32)
            IL = IL - IL[w, w]
            iniu, endu = suR0 + f32(L - w), suR0 + f32(L + w + 1)
            if iniu < 0 or endu >= pyrR[levelL].shape[1]:
                continue
            bestSAD, bestinc, vD = 2 ** 31 - 1, 0, [f32(0)] * (2 * L + 1)
            for inc in range(-L, L + 1):
                x0 = int(suR0) + inc - w
                IR = pyrR[levelL][int(svL) - w:int(svL) + w + 1, x0:x0 + 2 * w + 1].astype(f32)
                IR = IR - IR[w, w]
                dist = f32(cv2.norm(IL, IR, cv2.NORM_L1))
                if dist < bestSAD:
                    bestSAD, bestinc = int(dist), inc
                vD[L + inc] = dist
            if bestinc in (-L, L):
                continue
            d1, d2, d3 = vD[L + bestinc - 1], vD[L + bestinc], vD[L + bestinc + 1]
            with np.errstate(invalid="ignore", divide="ignore"):
                deltaR = (d1 - d3) / (f32(2.0) * (d1 + d3 - f32(2.0) * d2))
            if deltaR < -1 or deltaR > 1:
                continue
            bestuR = scale[levelL] * (suR0 + f32(bestinc) + deltaR)
            disp = uL - bestuR
            if minD <= disp < maxD:
                if disp <= 0:
                    disp = f32(0.01)
                    bestuR = f32(float(uL) - 0.01)
                depth[iL] = mbf / disp
                uRight[iL] = bestuR
                vDistIdx.append((bestSAD, iL))
    if vDistIdx:
        vDistIdx.sort()
        median = f32(vDistIdx[len(vDistIdx) // 2][0])
        th = f32(f32(1.5) * f32(1.4)) * median
        for sad, iL in reversed(vDistIdx):
            if f32(sad) < th:
                break
            uRight[iL] = -1
            depth[iL] = -1
    return uRight, depth
