// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/api/user/algorithm/preduce.h:
// sequential fold in lexicographic order; `.get()` yields the value.
#pragma once
#include <utility>
#include "allscale/api/user/algorithm/pfor.h"

namespace allscale { namespace api { namespace user { namespace algorithm {

	namespace detail {
		template<typename R> struct reduction_result { R value; R get() { return std::move(value); } };   // results may be move-only (pic_tracer's ParticleCount)
	}

	// range of grid points: fold(index, acc)
	template<typename E, std::size_t D, typename Fold, typename Reduce, typename Init>
	auto preduce(const allscale::utils::Vector<E,D>& lo, const allscale::utils::Vector<E,D>& hi, const Fold& fold, const Reduce&, const Init& init)
			-> detail::reduction_result<decltype(init())> {
		auto acc = init();
		detail::forEach(lo, hi, [&](const allscale::utils::Vector<E,D>& p) { fold(p, acc); });
		return { acc };
	}

	// container: fold(element, acc)
	template<typename C, typename Fold, typename Reduce, typename Init, typename = decltype(std::declval<const C&>().begin())>
	auto preduce(const C& c, const Fold& fold, const Reduce&, const Init& init) -> detail::reduction_result<decltype(init())> {
		auto acc = init();
		for (const auto& e : c) fold(e, acc);
		return { acc };
	}

	// integer range with map(a,b)->R and reduce(R,R): the form pic_tracer.cxx:272 uses
	template<typename I, typename Map, typename Reduce, typename = typename std::enable_if<std::is_integral<I>::value>::type>
	auto preduce(I lo, I hi, const Map& map, const Reduce&) -> detail::reduction_result<decltype(map(lo, hi))> {
		return { map(lo, hi) };
	}

}}}}
