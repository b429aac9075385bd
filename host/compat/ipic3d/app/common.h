// Source compatibility: the reference's ipic3d/app/common.h (an include hub, hdr/common.h:1-9).
#pragma once
#include "ipic3d/app/universe.h"
#include "allscale/api/user/algorithm/preduce.h"
