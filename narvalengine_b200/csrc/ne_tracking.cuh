// ne_tracking.cuh — GridMedia::Tr (ratio tracking, materials/GridMedia.cpp:45-69) and GridMedia::sample (delta
// tracking, :71-100) as RESUMABLE walks. Included by ne_device.cuh (needs density_at, BrickDDA, bsdf_sample).
//
// A Tracker produces the next candidate collision point of an OCS ray segment [0, tFar]:
//   BRICKMAJ = false   the reference's walk: t -= log(1 - xi) * invMaxDensity / sigma_bar with the single global majorant
//                      (draw for draw what GridMedia does; used by the tape tests and NE_B200_RENDER_GLOBAL_MAJORANT)
//   BRICKMAJ = true    the same exponential walk against the majorant of the 8^3 brick the point is in, re-started at
//                      every brick boundary (memoryless, so the free-flight distribution is unchanged); empty bricks
//                      are crossed without a sample
// `budget` bounds the events (brick moves + candidates) of one call so that a wavefront kernel can stop a long walk,
// move the path's origin to the point reached and queue the rest for its next pass — again exact by memorylessness.
#pragma once

namespace ne {

enum { TRACK_END = 0, TRACK_CANDIDATE = 1, TRACK_BUDGET = 2 };

template <bool BRICKMAJ>
struct Tracker {
	Ray ray;  // OCS, origin at the segment start
	float t, tFar, sig;
	float invMaj;  // 1 / current majorant density (global: GridMedia::invMaxDensity)
	float step;    // invMaj / sig
	float tExit;   // end of the current brick (BRICKMAJ)
	float maj;
	BrickDDA dda;

	NE_D void init(const DVolume& v, const DMaterial& m, Ray rayOCS, float tStart, float tEnd, Stats& st) {
		ray = rayOCS;
		t = tStart;
		tFar = tEnd;
		V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
		sig = avg(ext * m.density_mult);
		if (BRICKMAJ) {
			dda.init(v, ray);  // tStart is 0 for every brick walk (the origin has been moved to the segment start)
			enter_brick(v, st);
		} else {
			invMaj = v.inv_max_density;
		}
	}
	NE_D void enter_brick(const DVolume& v, Stats& st) {
		st.brick_visits++;
		tExit = fminf(dda.exit_t(), tFar);
		maj = dda.majorant(v);
		invMaj = 1.0f / maj;
		step = invMaj / sig;
	}
	// Advance to the next candidate. One uniform per exponential sample.
	template <class R>
	NE_D int next(const DVolume& v, R& rng, Stats& st, int& budget) {
		if (!BRICKMAJ) {
			if (budget-- <= 0) return TRACK_BUDGET;
			t -= logf(1 - rng.next()) * invMaj / sig;  // GridMedia.cpp:56 / :82
			return t >= tFar ? TRACK_END : TRACK_CANDIDATE;
		}
		while (true) {
			if (budget-- <= 0) return TRACK_BUDGET;
			if (maj > 0) {
				t -= logf(1 - rng.next()) * step;
				if (t < tExit) return TRACK_CANDIDATE;
			}
			t = tExit;
			if (tExit >= tFar) return TRACK_END;
			if (!dda.step(v)) return TRACK_END;
			enter_brick(v, st);
		}
	}
};

#define NE_NO_BUDGET 0x7fffffff

// Ratio tracking over the OCS segment with pbrt's Russian roulette (GridMedia.cpp:58-66). `Tr` carries the running
// transmittance in and out (1 at the start). Returns TRACK_END when finished (Tr final, possibly 0 = killed) or
// TRACK_BUDGET (ray origin should be moved to trk.t by the caller).
template <class R, bool BRICKMAJ>
NE_D int ratio_walk(const DVolume& v, Tracker<BRICKMAJ>& trk, float& Tr, R& rng, Stats& st, int budget) {
	while (true) {
		int e = trk.next(v, rng, st, budget);
		if (e != TRACK_CANDIDATE) return e;
		st.ratio_steps++;
		float density = density_at(v, trk.ray, trk.t);
		Tr *= 1 - fmaxf(0.0f, density * trk.invMaj);
		const float rrThreshold = .1f;
		if (Tr < rrThreshold) {
			float q = fmaxf(0.05f, 1.0f - Tr);
			if (rng.next() < q) { Tr = 0.0f; return TRACK_END; }
			Tr /= 1 - q;
		}
	}
}

// Delta tracking: TRACK_CANDIDATE = real collision at trk.t, TRACK_END = escaped, TRACK_BUDGET = stopped at trk.t.
template <class R, bool BRICKMAJ>
NE_D int delta_walk(const DVolume& v, Tracker<BRICKMAJ>& trk, R& rng, Stats& st, int budget) {
	while (true) {
		int e = trk.next(v, rng, st, budget);
		if (e != TRACK_CANDIDATE) return e;
		st.delta_steps++;
		float density = density_at(v, trk.ray, trk.t);
		float ra = rng.next();
		if (density * trk.invMaj > ra) return TRACK_CANDIDATE;
	}
}

// GridMedia::Tr. rayW: WCS ray; tNear/tFar from the hit record. Returns the scalar transmittance.
template <class R, bool BRICKMAJ>
NE_D float grid_tr(const DInstance& in, const DMaterial& m, const DVolume& v, Ray rayW, float tNear, float tFar, R& rng, Stats& st) {
	Ray ray = transform_ray(rayW, in.Mi);
	ray.o = ray.at(tNear);
	Tracker<BRICKMAJ> trk;
	trk.init(v, m, ray, 0.0f, tFar - tNear, st);
	float Tr = 1;
	ratio_walk<R, BRICKMAJ>(v, trk, Tr, rng, st, NE_NO_BUDGET);
	return Tr;
}

// The scattering vertex of GridMedia::sample (:90-95) at parameter t of the OCS ray: phase sample about +Y of the
// OCS (Q19: ignores the incoming direction), mapped to WCS by M (the direction keeps the instance scale).
template <class R>
NE_D Ray grid_scatter(const DScene& s, const DInstance& in, const DMaterial& m, const Ray& rayOCS, float t, const Hit& isect, R& rng, Stats& st) {
	st.scatter_events++;
	Ray so;
	so.o = rayOCS.at(t);
	so.d = bsdf_sample(s, m, rayOCS.d, V3(0.0f, 1.0f, 0.0f), isect, rng);
	return transform_ray(so, in.M);
}

// GridMedia::sample. In Li `incoming` already has its origin at the segment start and tNear = 0.
// Returns the value the reference returns: sigma_s/sigma_t on a collision, exactly (1,1,1) on escape (Q1, Q1b).
template <class R, bool BRICKMAJ>
NE_D V3 grid_sample(const DScene& s, const DInstance& in, const DMaterial& m, const DVolume& v, Ray incomingW, float tNear, float tFar,
                    const Hit& isect, Ray& scattered, R& rng, Stats& st) {
	scattered = incomingW;
	Ray ray = transform_ray(incomingW, in.Mi);
	Tracker<BRICKMAJ> trk;
	trk.init(v, m, ray, tNear, tFar, st);  // GridMedia.cpp:79 starts at tNear; Li always passes 0
	if (delta_walk<R, BRICKMAJ>(v, trk, rng, st, NE_NO_BUDGET) != TRACK_CANDIDATE) return V3(1.0f);
	scattered = grid_scatter(s, in, m, ray, trk.t, isect, rng, st);
	V3 sc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
	V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + sc;
	return sc / ext;
}

}  // namespace ne
