// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/api/user/data/grid.h:
// GridPoint<D> (= int64 vector) and a dense, value-initialised, movable Grid<T,D>
// addressed by GridPoint; row-major with the last coordinate fastest (matches the
// reference's own flatten, cell.h:521-523).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include "allscale/utils/vector.h"
#include "allscale/utils/serializer.h"

namespace allscale { namespace api { namespace user { namespace data {

	template<std::size_t D> using GridPoint = allscale::utils::Vector<std::int64_t, D>;

	template<typename T, std::size_t D = 3>
	class Grid {
		GridPoint<D> extent;
		std::vector<T> cells;
		std::size_t flat(const GridPoint<D>& p) const {
			std::size_t idx = 0;
			for (std::size_t i = 0; i < D; ++i) idx = idx * (std::size_t)extent[i] + (std::size_t)p[i];
			return idx;
		}
		static std::size_t volume(const GridPoint<D>& s) { std::size_t v = 1; for (std::size_t i = 0; i < D; ++i) v *= (std::size_t)s[i]; return v; }
	public:
		using element_type = T;
		using coordinate_type = GridPoint<D>;

		Grid() : extent(0) {}
		Grid(const coordinate_type& size) : extent(size), cells(volume(size)) {}
		Grid(const Grid&) = delete;
		Grid& operator=(const Grid&) = delete;
		Grid(Grid&&) = default;
		Grid& operator=(Grid&&) = default;

		coordinate_type size() const { return extent; }
		T& operator[](const coordinate_type& p) { return cells[flat(p)]; }
		const T& operator[](const coordinate_type& p) const { return cells[flat(p)]; }

		template<typename Op> void forEach(const Op& op) { for (auto& c : cells) op(c); }
		template<typename Op> void forEach(const Op& op) const { for (const auto& c : cells) op(c); }
	};

}}}}
