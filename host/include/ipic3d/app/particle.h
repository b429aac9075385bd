// Host C++ mirror: `Particle` (reference hdr/particle.h:9-49).  Same 64-byte POD the C ABI transfers
// (include/ipic3d_b200.h "particles"), plus the scalar Boris step for host-side callers and tests; the
// per-timestep work itself runs in the CUDA kernels (csrc/ipic_device.cuh boris_push).
#pragma once
#include "ipic3d/app/vector.h"

namespace ipic3d {

	struct Particle {
		Vector3<double> position;
		Vector3<double> velocity;
		double q = 0;       // charge
		double qom = 0;     // charge over mass

		// x += v * dt
		void updatePosition(double dt) { position += velocity * dt; }

		// Boris rotation with half-step electric kicks (reference hdr/particle.h:31-49):
		//   k = qom/2 * dt;  t = k B;  s = 2 t / (1 + |t|^2);  v- = v + k E;  v' = v- + v- x t;  v+ = v- + v' x s;  v = v+ + k E
		void updateVelocity(const Vector3<double>& E, const Vector3<double>& B, double dt) {
			const double k = qom * 0.5 * dt;
			const Vector3<double> t = k * B;
			const double t2 = sumOfSquares(t);
			const Vector3<double> s = (2.0 * t) / (1.0 + t2);
			const Vector3<double> halfKick = k * E;
			const Vector3<double> vMinus = velocity + halfKick;
			const Vector3<double> vPrime = vMinus + crossProduct(vMinus, t);
			const Vector3<double> vPlus = vMinus + crossProduct(vPrime, s);
			velocity = vPlus + halfKick;
		}
	};
	static_assert(sizeof(Particle) == 64, "Particle must match the 8-double AoS record of the C ABI");

} // namespace ipic3d
