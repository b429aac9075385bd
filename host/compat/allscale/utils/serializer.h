// Source compatibility: `allscale::utils::is_serializable<T>` (AllScale API, utils/serializer.h).  AllScale serialises
// objects to move them between the localities of its runtime.  The host mirror has one address space per process and
// hands particle data to the GPUs as plain FP64 records through the C ABI, so the serialised form of an object here is
// its bytes: a type is serialisable when a byte copy reproduces it.
#pragma once
#include <type_traits>

namespace allscale { namespace utils {
	template<typename T>
	struct is_serializable : std::is_trivially_copyable<T> {};
}}
