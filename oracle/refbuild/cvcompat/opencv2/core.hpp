// OpenCV-compat layer (oracle/refbuild): see opencv2/core/core.hpp.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "core/core.hpp"
