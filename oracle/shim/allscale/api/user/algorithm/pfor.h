// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/api/user/algorithm/pfor.h.
// pfor(begin,end,body) / pfor(extent,body) over GridPoint ranges, executed before returning
// (the real API returns a handle whose destructor is a barrier; every reference call site
// lets the handle die at once, so "run then return" has the same observable semantics).
// Default: sequential, lexicographic order (first coordinate slowest).
// With -DALLSCALE_SHIM_THREADS the iteration space is spread over std::thread workers (count from
// env NUM_WORKERS, the reference's own knob, else hardware_concurrency) pulling chunks off an
// atomic counter; the reference's loop bodies touch disjoint data per index, so
// both variants give identical results.  (No OpenMP: this image's gcc ships no libgomp.spec.)
#pragma once
#include <cstdint>
#ifdef ALLSCALE_SHIM_THREADS
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>
#endif
#include "allscale/utils/vector.h"

namespace allscale { namespace api { namespace user { namespace algorithm {

	namespace detail {

		struct loop_reference { void wait() const {} };

#ifdef ALLSCALE_SHIM_THREADS
		inline int worker_count() {
			static const int n = [] {
				if (const char* e = std::getenv("NUM_WORKERS")) { int v = std::atoi(e); if (v > 0) return v; }
				unsigned h = std::thread::hardware_concurrency();
				return h ? (int)h : 1;
			}();
			return n;
		}
#endif

		// sequential N-d iteration, first coordinate slowest
		template<typename E, std::size_t D, typename Body>
		void forEach(const allscale::utils::Vector<E,D>& lo, const allscale::utils::Vector<E,D>& hi, const Body& body) {
			for (std::size_t i = 0; i < D; ++i) if (!(lo[i] < hi[i])) return;
			allscale::utils::Vector<E,D> cur(lo);
			for (;;) {
				body(cur);
				std::size_t d = D;
				while (d > 0) {
					--d;
					if (++cur[d] < hi[d]) break;
					cur[d] = lo[d];
					if (d == 0) return;
				}
			}
		}
	}

	template<typename E, std::size_t D, typename Body>
	detail::loop_reference pfor(const allscale::utils::Vector<E,D>& lo, const allscale::utils::Vector<E,D>& hi, const Body& body) {
#ifdef ALLSCALE_SHIM_THREADS
		for (std::size_t i = 0; i < D; ++i) if (!(lo[i] < hi[i])) return {};
		std::int64_t total = 1;
		for (std::size_t i = 0; i < D; ++i) total *= (std::int64_t)(hi[i] - lo[i]);
		const int workers = detail::worker_count();
		if (workers <= 1 || total < 64) { detail::forEach(lo, hi, body); return {}; }
		const std::int64_t chunk = 16;
		std::atomic<std::int64_t> next(0);
		auto work = [&]() {
			for (;;) {
				std::int64_t b = next.fetch_add(chunk);
				if (b >= total) return;
				std::int64_t e = b + chunk < total ? b + chunk : total;
				for (std::int64_t lin = b; lin < e; ++lin) {
					allscale::utils::Vector<E,D> cur; std::int64_t rest = lin;
					for (std::size_t d = D; d-- > 0;) { std::int64_t ext = (std::int64_t)(hi[d] - lo[d]); cur[d] = lo[d] + (E)(rest % ext); rest /= ext; }
					body(cur);
				}
			}
		};
		std::vector<std::thread> pool;
		for (int t = 1; t < workers; ++t) pool.emplace_back(work);
		work();
		for (auto& t : pool) t.join();
#else
		detail::forEach(lo, hi, body);
#endif
		return {};
	}

	template<typename E, std::size_t D, typename Body>
	detail::loop_reference pfor(const allscale::utils::Vector<E,D>& extent, const Body& body) {
		return pfor(allscale::utils::Vector<E,D>(E(0)), extent, body);
	}

	// scalar ranges
	template<typename I, typename Body, typename = typename std::enable_if<std::is_integral<I>::value>::type>
	detail::loop_reference pfor(I lo, I hi, const Body& body) { for (I i = lo; i < hi; ++i) body(i); return {}; }

	// containers
	template<typename C, typename Body, typename = decltype(std::declval<C&>().begin())>
	detail::loop_reference pfor(C& c, const Body& body) { for (auto& e : c) body(e); return {}; }

}}}}
