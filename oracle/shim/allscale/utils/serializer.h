// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/utils/serializer.h: byte archive
// with write<T>/read<T>, the trivially_serializable tag and the is_serializable trait.
#pragma once
#include <cstring>
#include <type_traits>
#include <vector>

namespace allscale { namespace utils {

	struct trivially_serializable {};

	class ArchiveWriter;
	class ArchiveReader;

	namespace shim_detail {
		template<typename T, typename = void> struct has_store_load : std::false_type {};
		template<typename T> struct has_store_load<T, decltype((void)std::declval<const T&>().store(std::declval<ArchiveWriter&>()), (void)T::load(std::declval<ArchiveReader&>()))> : std::true_type {};
	}

	template<typename T> struct is_trivially_serializable
		: std::integral_constant<bool, std::is_arithmetic<T>::value || std::is_enum<T>::value || std::is_base_of<trivially_serializable,T>::value> {};

	template<typename T> struct is_serializable
		: std::integral_constant<bool, is_trivially_serializable<T>::value || shim_detail::has_store_load<T>::value || std::is_trivially_copyable<T>::value> {};

	class ArchiveWriter {
		std::vector<char> bytes;
	public:
		template<typename T> typename std::enable_if<!shim_detail::has_store_load<T>::value>::type write(const T& v) {
			const char* p = reinterpret_cast<const char*>(&v); bytes.insert(bytes.end(), p, p + sizeof(T));
		}
		template<typename T> typename std::enable_if<shim_detail::has_store_load<T>::value>::type write(const T& v) { v.store(*this); }
		const std::vector<char>& data() const { return bytes; }
	};

	class ArchiveReader {
		const char* cur;
	public:
		explicit ArchiveReader(const std::vector<char>& b) : cur(b.data()) {}
		template<typename T> typename std::enable_if<!shim_detail::has_store_load<T>::value, T>::type read() {
			typename std::aligned_storage<sizeof(T), alignof(T)>::type raw; std::memcpy(&raw, cur, sizeof(T)); cur += sizeof(T);
			return *reinterpret_cast<T*>(&raw);
		}
		template<typename T> typename std::enable_if<shim_detail::has_store_load<T>::value, T>::type read() { return T::load(*this); }
	};

}}
