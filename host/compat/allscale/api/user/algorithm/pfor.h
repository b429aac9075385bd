// Source compatibility for code written against the reference: the AllScale API's `pfor` (external to the reference,
// fetched by its CMake) under its own name, implemented by the host mirror's thread-pool loop (ipic3d/app/vector.h).
// Only what the reference's own unit tests use (code/app/test/*.cc): pfor(begin, end, body) and pfor(extent, body)
// over 3-D coordinates.
#pragma once
#include "ipic3d/app/vector.h"

namespace allscale { namespace api { namespace user { namespace algorithm {
	template<typename Body>
	void pfor(const ipic3d::coordinate_type& begin, const ipic3d::coordinate_type& end, const Body& body) { ipic3d::pfor(begin, end, body); }
	template<typename Body>
	void pfor(const ipic3d::coordinate_type& extent, const Body& body) { ipic3d::pfor(ipic3d::coordinate_type(0), extent, body); }
}}}}
