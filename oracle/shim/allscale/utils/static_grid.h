// TEST INFRASTRUCTURE ONLY (oracle/). allscale/utils/static_grid.h is included by the
// reference (cell.h:12) but nothing from it is used; intentionally empty.
#pragma once
