// Source compatibility: the reference's ipic3d/app/common.h includes this AllScale header; nothing of it is used by the
// code that is compiled against the host mirror.
#pragma once
