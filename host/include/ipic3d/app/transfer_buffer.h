// Host C++ mirror: `TransferDirection` / `TransferBuffers` (reference hdr/transfer_buffer.h:12-96), the 27 particle
// lists every cell exports into.  During `simulateSteps` their role is played on the GPU by the tagged outbox at the
// tail of each cell segment and its direction-sorted index (csrc/particle_kernels.cu); this host container keeps the
// reference's type for code that names it (mover functor signatures, the reference's transfer_buffer_test.cc) and for
// staging particles on the host.  One flat array of lists, 27 consecutive lists per cell.
#pragma once
#include <cstddef>
#include <vector>

#include "ipic3d/app/particle.h"
#include "ipic3d/app/vector.h"

namespace ipic3d {

	// one of the 3 x 3 x 3 neighbours: per axis 0 = predecessor, 1 = same plane, 2 = successor
	class TransferDirection {
		friend class TransferBuffers;
		unsigned slot;                     // 9 x + 3 y + z, the reference's traversal order of the 27 buffers
	public:
		using Direction = int;
		constexpr static Direction Predecessor = 0;
		constexpr static Direction Center = 1;
		constexpr static Direction Successor = 2;
		TransferDirection(Direction x, Direction y, Direction z) : slot((unsigned)(9 * x + 3 * y + z)) {}

	};

	class TransferBuffers {
		coordinate_type extent;
		std::vector<std::vector<Particle>> lists;       // [flatten(cell) * 27 + direction]
	public:
		using grid_size_t = coordinate_type;
		using grid_pos_t = coordinate_type;

		explicit TransferBuffers(const grid_size_t& size) : extent(size), lists((std::size_t)(size.x * size.y * size.z) * 27) {}

		std::vector<Particle>& getBuffer(const grid_pos_t& pos, const TransferDirection& dir) {
			return lists[(std::size_t)((pos.x * extent.y + pos.y) * extent.z + pos.z) * 27 + dir.slot];
		}
		const std::vector<Particle>& getBuffer(const grid_pos_t& pos, const TransferDirection& dir) const {
			return lists[(std::size_t)((pos.x * extent.y + pos.y) * extent.z + pos.z) * 27 + dir.slot];
		}
		const grid_size_t& size() const { return extent; }
	};

} // namespace ipic3d
