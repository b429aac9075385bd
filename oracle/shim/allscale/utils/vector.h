// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/utils/vector.h of the un-vendored
// AllScale API (pinned by /root/reference/cmake/allscale.cmake:42-54).  Restated from the
// call sites in the reference (grep 'allscale::utils' code/app) and the published semantics:
// fixed-size algebraic vector, element-wise arithmetic, lexicographic '<', dominance tests,
// "[a,b,c]" printing.  Floating-point evaluation order follows the natural left-to-right
// component loops; nothing in the reference's tests pins this at bit level.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <ostream>
#include <type_traits>
#include "allscale/utils/assert.h"

namespace allscale { namespace utils {

	namespace shim_detail {
		// storage: N==3 exposes x,y,z; N==2 exposes x,y; otherwise an array
		template<typename T, std::size_t N> struct VecStore {
			std::array<T,N> data{};
			T& at(std::size_t i) { return data[i]; }
			const T& at(std::size_t i) const { return data[i]; }
		};
		template<typename T> struct VecStore<T,3> {
			T x{}, y{}, z{};
			T& at(std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
			const T& at(std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
		};
		template<typename T> struct VecStore<T,2> {
			T x{}, y{};
			T& at(std::size_t i) { return i == 0 ? x : y; }
			const T& at(std::size_t i) const { return i == 0 ? x : y; }
		};
	}

	template<typename T, std::size_t N>
	class Vector : public shim_detail::VecStore<T,N> {
		using store = shim_detail::VecStore<T,N>;
	public:
		using element_type = T;

		Vector() = default;
		Vector(const Vector&) = default;
		Vector& operator=(const Vector&) = default;

		// broadcast one scalar to all components (implicit on purpose: `Size<3> s = 1;`)
		Vector(const T& e) { for (std::size_t i = 0; i < N; ++i) store::at(i) = e; }

		template<std::size_t M = N, typename = typename std::enable_if<M == 3>::type>
		Vector(const T& a, const T& b, const T& c) { store::at(0) = a; store::at(1) = b; store::at(2) = c; }

		template<std::size_t M = N, typename = typename std::enable_if<M == 2>::type>
		Vector(const T& a, const T& b) { store::at(0) = a; store::at(1) = b; }

		// element-type conversion
		template<typename R, typename = typename std::enable_if<!std::is_same<R,T>::value>::type>
		Vector(const Vector<R,N>& o) { for (std::size_t i = 0; i < N; ++i) store::at(i) = (T)o[i]; }

		T& operator[](std::size_t i) { return store::at(i); }
		const T& operator[](std::size_t i) const { return store::at(i); }

		// -- compound assignment --
		Vector& operator+=(const Vector& o) { for (std::size_t i = 0; i < N; ++i) store::at(i) += o[i]; return *this; }
		Vector& operator-=(const Vector& o) { for (std::size_t i = 0; i < N; ++i) store::at(i) -= o[i]; return *this; }
		template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
		Vector& operator*=(const S& f) { for (std::size_t i = 0; i < N; ++i) store::at(i) *= f; return *this; }
		template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
		Vector& operator/=(const S& f) { for (std::size_t i = 0; i < N; ++i) store::at(i) /= f; return *this; }

		// -- comparison --
		bool operator==(const Vector& o) const { for (std::size_t i = 0; i < N; ++i) if (!(store::at(i) == o[i])) return false; return true; }
		bool operator!=(const Vector& o) const { return !(*this == o); }
		// lexicographic order (used by the field solvers' ghost guard, reference field.h:315,366)
		bool operator<(const Vector& o) const {
			for (std::size_t i = 0; i < N; ++i) { if (store::at(i) < o[i]) return true; if (o[i] < store::at(i)) return false; }
			return false;
		}
		bool operator<=(const Vector& o) const { return !(o < *this); }
		bool operator>(const Vector& o) const { return o < *this; }
		bool operator>=(const Vector& o) const { return !(*this < o); }
		bool dominatedBy(const Vector& o) const { for (std::size_t i = 0; i < N; ++i) if (!(store::at(i) <= o[i])) return false; return true; }
		bool strictlyDominatedBy(const Vector& o) const { for (std::size_t i = 0; i < N; ++i) if (!(store::at(i) < o[i])) return false; return true; }

		friend std::ostream& operator<<(std::ostream& out, const Vector& v) {
			out << "[";
			for (std::size_t i = 0; i < N; ++i) { if (i) out << ","; out << v[i]; }
			return out << "]";
		}
	};

	// -- binary arithmetic (vector ∘ vector) --
	template<typename T, std::size_t N> Vector<T,N> operator+(const Vector<T,N>& a, const Vector<T,N>& b) { Vector<T,N> r(a); r += b; return r; }
	template<typename T, std::size_t N> Vector<T,N> operator-(const Vector<T,N>& a, const Vector<T,N>& b) { Vector<T,N> r(a); r -= b; return r; }
	template<typename T, std::size_t N> Vector<T,N> operator-(const Vector<T,N>& a) { Vector<T,N> r; for (std::size_t i = 0; i < N; ++i) r[i] = -a[i]; return r; }
	template<typename T, std::size_t N> Vector<T,N> operator%(const Vector<T,N>& a, const Vector<T,N>& b) { Vector<T,N> r; for (std::size_t i = 0; i < N; ++i) r[i] = a[i] % b[i]; return r; }

	// -- vector ∘ scalar --
	template<typename T, std::size_t N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
	Vector<T,N> operator*(const Vector<T,N>& a, const S& f) { Vector<T,N> r(a); r *= f; return r; }
	template<typename T, std::size_t N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
	Vector<T,N> operator*(const S& f, const Vector<T,N>& a) { Vector<T,N> r; for (std::size_t i = 0; i < N; ++i) r[i] = f * a[i]; return r; }
	template<typename T, std::size_t N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
	Vector<T,N> operator/(const Vector<T,N>& a, const S& f) { Vector<T,N> r(a); r /= f; return r; }

	// -- free algebra helpers --
	template<typename T, std::size_t N>
	T sumOfSquares(const Vector<T,N>& v) { T s = T(); for (std::size_t i = 0; i < N; ++i) s += v[i] * v[i]; return s; }

	template<typename T>
	Vector<T,3> crossProduct(const Vector<T,3>& a, const Vector<T,3>& b) {
		return Vector<T,3>(a[1]*b[2] - a[2]*b[1], a[2]*b[0] - a[0]*b[2], a[0]*b[1] - a[1]*b[0]);
	}

	template<typename A, typename B, std::size_t N>
	Vector<decltype(A() * B()),N> elementwiseProduct(const Vector<A,N>& a, const Vector<B,N>& b) {
		Vector<decltype(A() * B()),N> r; for (std::size_t i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r;
	}
	template<typename A, typename B, std::size_t N>
	Vector<decltype(A() / B()),N> elementwiseDivision(const Vector<A,N>& a, const Vector<B,N>& b) {
		Vector<decltype(A() / B()),N> r; for (std::size_t i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r;
	}

}}
