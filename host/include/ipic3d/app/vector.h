// Host C++ mirror of the reference's data model — vector and grid types.
//
// The reference builds on the AllScale API (`allscale::utils::Vector<T,3>`, `GridPoint<3>`, `Grid<T,3>`;
// reference hdr/vector.h:1-16, allscale_api pinned in cmake/allscale.cmake:42-54).  The B200 build has no
// AllScale runtime underneath (its role is taken by the C ABI in include/ipic3d_b200.h), so the same
// surface is provided here as plain host containers: element-wise arithmetic, dominance tests,
// "[a,b,c]" printing, a dense value-initialised 3-D grid flattened (x*Ny + y)*Nz + z.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <ostream>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace ipic3d {

	template<typename T>
	struct Vector3 {
		T x{}, y{}, z{};

		constexpr Vector3() = default;
		constexpr Vector3(T all) : x(all), y(all), z(all) {}                     // broadcast, implicit like the AllScale type
		constexpr Vector3(T a, T b, T c) : x(a), y(b), z(c) {}
		template<typename R, typename = typename std::enable_if<!std::is_same<R, T>::value>::type>
		constexpr Vector3(const Vector3<R>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}

		T& operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
		const T& operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }

		Vector3& operator+=(const Vector3& o) { x += o.x; y += o.y; z += o.z; return *this; }
		Vector3& operator-=(const Vector3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
		template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
		Vector3& operator*=(S f) { x *= f; y *= f; z *= f; return *this; }
		template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
		Vector3& operator/=(S f) { x /= f; y /= f; z /= f; return *this; }

		bool operator==(const Vector3& o) const { return x == o.x && y == o.y && z == o.z; }
		bool operator!=(const Vector3& o) const { return !(*this == o); }
		// lexicographic, as the reference's field solvers rely on (hdr/field.h:315, :366)
		bool operator<(const Vector3& o) const {
			if (x != o.x) return x < o.x;
			if (y != o.y) return y < o.y;
			return z < o.z;
		}
		bool dominatedBy(const Vector3& o) const { return x <= o.x && y <= o.y && z <= o.z; }
		bool strictlyDominatedBy(const Vector3& o) const { return x < o.x && y < o.y && z < o.z; }

		friend std::ostream& operator<<(std::ostream& out, const Vector3& v) { return out << "[" << v.x << "," << v.y << "," << v.z << "]"; }
	};

	template<typename T> Vector3<T> operator+(Vector3<T> a, const Vector3<T>& b) { return a += b; }
	template<typename T> Vector3<T> operator-(Vector3<T> a, const Vector3<T>& b) { return a -= b; }
	template<typename T> Vector3<T> operator-(const Vector3<T>& a) { return { -a.x, -a.y, -a.z }; }
	template<typename T> Vector3<T> operator%(const Vector3<T>& a, const Vector3<T>& b) { return { a.x % b.x, a.y % b.y, a.z % b.z }; }
	template<typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
	Vector3<T> operator*(Vector3<T> a, S f) { return a *= f; }
	template<typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
	Vector3<T> operator*(S f, const Vector3<T>& a) { return { (T)(f * a.x), (T)(f * a.y), (T)(f * a.z) }; }
	template<typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
	Vector3<T> operator/(Vector3<T> a, S f) { return a /= f; }

	template<typename T> T sumOfSquares(const Vector3<T>& v) { T s = T(); s += v.x * v.x; s += v.y * v.y; s += v.z * v.z; return s; }
	template<typename T> T norm(const Vector3<T>& v) { return std::sqrt(sumOfSquares(v)); }        // reference hdr/vector.h:12-15
	template<typename T> Vector3<T> crossProduct(const Vector3<T>& a, const Vector3<T>& b) {
		return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
	}
	template<typename A, typename B> Vector3<decltype(A() * B())> elementwiseProduct(const Vector3<A>& a, const Vector3<B>& b) {
		return { a.x * b.x, a.y * b.y, a.z * b.z };
	}
	template<typename A, typename B> Vector3<decltype(A() / B())> elementwiseDivision(const Vector3<A>& a, const Vector3<B>& b) {
		return { a.x / b.x, a.y / b.y, a.z / b.z };
	}

	// grid coordinates / sizes (the reference's allscale GridPoint<3>, utils::Coordinate<3>, utils::Size<3>)
	using coordinate_type = Vector3<std::int64_t>;
	namespace utils {
		template<std::size_t D> using Coordinate = coordinate_type;     // only D == 3 exists on this path
		template<std::size_t D> using Size = coordinate_type;
		template<typename T, std::size_t D> using vec = Vector3<T>;
	}

	// dense, value-initialised 3-D container (the reference's allscale Grid<T,3>); z runs fastest
	template<typename T>
	class Grid3 {
		coordinate_type extent;
		std::vector<T> cells;
	public:
		explicit Grid3(const coordinate_type& size = coordinate_type(0))
			: extent(size), cells((std::size_t)(size.x * size.y * size.z)) {}
		Grid3(Grid3&&) = default;
		Grid3& operator=(Grid3&&) = default;
		Grid3(const Grid3&) = delete;                                       // like the AllScale grid: movable only
		Grid3& operator=(const Grid3&) = delete;

		const coordinate_type& size() const { return extent; }
		std::size_t flatten(const coordinate_type& p) const { return (std::size_t)((p.x * extent.y + p.y) * extent.z + p.z); }
		T& operator[](const coordinate_type& p) { return cells[flatten(p)]; }
		const T& operator[](const coordinate_type& p) const { return cells[flatten(p)]; }
		T* data() { return cells.data(); }
		const T* data() const { return cells.data(); }
		std::size_t count() const { return cells.size(); }

		template<typename Body> void forEach(const Body& body) { for (auto& c : cells) body(c); }
	};

	// sequential counterpart of the reference's `pfor(begin, end, body)` / `forEach` loops, lexicographic order
	template<typename Body>
	void forEachCoordinate(const coordinate_type& begin, const coordinate_type& end, const Body& body) {
		for (std::int64_t i = begin.x; i < end.x; ++i)
			for (std::int64_t j = begin.y; j < end.y; ++j)
				for (std::int64_t k = begin.z; k < end.z; ++k) body(coordinate_type{ i, j, k });
	}

	// `pfor` proper: the body is called once per coordinate from a pool of host threads, in no particular order.
	// Only for bodies that touch nothing but their own coordinate's data (per-cell initialisation: every cell has
	// its own seed, so the result does not depend on the schedule).  NUM_WORKERS caps the threads like in the
	// reference's runs.
	template<typename Body>
	void pfor(const coordinate_type& begin, const coordinate_type& end, const Body& body) {
		const std::int64_t ny = end.y - begin.y, nz = end.z - begin.z;
		const std::int64_t total = (end.x - begin.x) * ny * nz;
		if (total <= 0) return;
		unsigned workers = std::thread::hardware_concurrency();
		if (const char* e = std::getenv("NUM_WORKERS")) workers = (unsigned)std::atoi(e);
		workers = std::max(1u, std::min(workers, 64u));
		if (total < 64 || workers == 1) { forEachCoordinate(begin, end, body); return; }
		std::atomic<std::int64_t> next(0);
		auto work = [&]() {
			for (;;) {
				const std::int64_t t = next.fetch_add(1);
				if (t >= total) return;
				body(coordinate_type{ begin.x + t / (ny * nz), begin.y + (t / nz) % ny, begin.z + t % nz });
			}
		};
		std::vector<std::thread> pool;
		for (unsigned w = 1; w < workers; ++w) pool.emplace_back(work);
		work();
		for (auto& t : pool) t.join();
	}

} // namespace ipic3d
