#!/usr/bin/env python3
"""Cut the hot-path functions out of the reference's Frame.cc / ORBmatcher.cc, VERBATIM, into include fragments.

TEST INFRASTRUCTURE ONLY (part of the oracle/_ref recipe, see build.sh).  Frame.cc and ORBmatcher.cc as whole files need
Eigen, DBoW2, g2o, glog, gflags and boost, none of which exist in this image; the three pieces on the hot path need none
of them, so they are copied byte-for-byte — no edits — into oracle/_ref/gen/ (git-ignored: reference sources never enter
this repository) and compiled inside frame_stub.cpp against a stub `Frame` that only declares the members they touch.
ORBextractor.cc needs no extraction: it is compiled where it lies.

Each fragment is located by line range AND checked against its expected first/last lines, so a reference checkout that
moved the code fails loudly instead of compiling something else.
"""
import os
import sys

FRAGMENTS = [
    # (source file, first line, last line, must start with, must end with, output name)
    ("src/Frame.cc", 758, 932, "void Frame::ComputeStereoMatches()", "}", "Frame_ComputeStereoMatches.inc"),
    ("src/ORBmatcher.cc", 37, 39, "const int ORBmatcher::TH_HIGH = 100;", "const int ORBmatcher::HISTO_LENGTH = 30;", "ORBmatcher_constants.inc"),
    ("src/ORBmatcher.cc", 1698, 1716, "// Bit set count operation from", "}", "ORBmatcher_DescriptorDistance.inc"),
]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/introspective_ORB_SLAM"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "_ref", "gen")
    os.makedirs(out, exist_ok=True)
    for rel, a, b, head, tail, name in FRAGMENTS:
        with open(os.path.join(ref, rel), "rb") as f:
            lines = f.read().split(b"\n")
        frag = lines[a - 1:b]
        first, last = frag[0].decode().strip(), frag[-1].decode().strip()
        if first != head or last != tail:
            sys.exit("%s:%d-%d does not hold the expected code (got %r ... %r)" % (rel, a, b, first, last))
        with open(os.path.join(out, name), "wb") as f:
            f.write(b"\n".join(frag) + b"\n")
        print("extracted %s:%d-%d -> %s" % (rel, a, b, name))


if __name__ == "__main__":
    main()
