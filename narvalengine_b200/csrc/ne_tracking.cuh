// ne_tracking.cuh — GridMedia::Tr (ratio tracking, materials/GridMedia.cpp:45-69) and GridMedia::sample (delta
// tracking, :71-100) as RESUMABLE walks made of single EVENTS. Included by ne_device.cuh (needs density_at,
// BrickDDA, bsdf_sample).
//
// A Tracker walks the OCS ray segment [0, tFar]; one event() either proposes a candidate collision point or moves
// to the next brick:
//   BRICKMAJ = false   the reference's walk: t -= log(1 - xi) * invMaxDensity / sigma_bar with the single global majorant
//                      (draw for draw what GridMedia does; used by the tape tests and NE_B200_RENDER_GLOBAL_MAJORANT)
//   BRICKMAJ = true    the same exponential walk against the majorant of the 8^3 brick the point is in, re-started at
//                      every brick boundary (memoryless, so the free-flight distribution is unchanged); empty bricks
//                      are crossed without a sample
// Because walks are sequences of independent events they can be stopped anywhere: the wavefront's tracking kernels
// keep all lanes of a warp busy by handing a lane the next queued walk as soon as its own ends, and bound the events
// of one pass (`budget`); a stopped walk moves its origin to the point reached and continues in the next pass.
#pragma once

namespace ne {

enum { TRACK_END = 0, TRACK_CANDIDATE = 1, TRACK_BUDGET = 2, TRACK_MOVED = 3 };

template <bool BRICKMAJ>
struct Tracker {
	Ray ray;  // OCS, origin at the segment start
	float t, tFar;
	float invSig;  // 1 / sigma_bar-per-unit-density = 1 / avg(extinction * densityMultiplier)
	float sig;
	float invMaj;  // 1 / current majorant density (global: GridMedia::invMaxDensity); 0 = empty brick
	float step;    // invMaj / sig: mean free path against the current majorant
	float tExit;   // end of the current brick (BRICKMAJ)
	BrickDDA dda;

	NE_D void init(const DVolume& v, const DMaterial& m, Ray rayOCS, float tStart, float tEnd, Stats& st) {
		ray = rayOCS;
		t = tStart;
		tFar = tEnd;
		V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
		sig = avg(ext * m.density_mult);
		if (BRICKMAJ) {
			invSig = 1.0f / sig;
			dda.init(v, ray);  // tStart is 0 for every brick walk (the origin has been moved to the segment start)
			enter_brick(v, st);
		} else {
			invMaj = v.inv_max_density;
		}
	}
	NE_D void enter_brick(const DVolume& v, Stats& st) {
		st.brick_visits++;
		tExit = fminf(dda.exit_t(), tFar);
		invMaj = dda.inv_majorant(v);
		step = invMaj * invSig;
	}
	// One event. TRACK_CANDIDATE: look the density up at `t`; TRACK_MOVED: entered the next brick; TRACK_END: the
	// segment is finished. One uniform per exponential sample.
	template <class R>
	NE_D int event(const DVolume& v, R& rng, Stats& st) {
		if (!BRICKMAJ) {
			t -= logf(1 - rng.next()) * invMaj / sig;  // GridMedia.cpp:56 / :82
			return t >= tFar ? TRACK_END : TRACK_CANDIDATE;
		}
		if (invMaj > 0) {
			rng.begin_event();
			t -= logf(1 - rng.next()) * step;
			if (t < tExit) return TRACK_CANDIDATE;
		}
		t = tExit;
		if (tExit >= tFar) return TRACK_END;
		if (!dda.step(v)) return TRACK_END;
		enter_brick(v, st);
		return TRACK_MOVED;
	}
};

#define NE_NO_BUDGET 0x7fffffff

// One ratio-tracking event with pbrt's Russian roulette (GridMedia.cpp:58-66). Returns TRACK_END when the walk is
// over (Tr final, possibly 0 = killed), otherwise TRACK_MOVED / TRACK_CANDIDATE (keep going).
template <class R, bool BRICKMAJ>
NE_D int ratio_event(const DVolume& v, Tracker<BRICKMAJ>& trk, float& Tr, R& rng, Stats& st) {
	int e = trk.event(v, rng, st);
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
	return TRACK_MOVED;
}
// One delta-tracking event (GridMedia.cpp:80-95). TRACK_CANDIDATE = REAL collision at trk.t, TRACK_END = escaped,
// TRACK_MOVED = keep going.
template <class R, bool BRICKMAJ>
NE_D int delta_event(const DVolume& v, Tracker<BRICKMAJ>& trk, R& rng, Stats& st) {
	int e = trk.event(v, rng, st);
	if (e != TRACK_CANDIDATE) return e;
	st.delta_steps++;
	float density = density_at(v, trk.ray, trk.t);
	float ra = rng.next();
	return density * trk.invMaj > ra ? TRACK_CANDIDATE : TRACK_MOVED;
}

template <class R, bool BRICKMAJ>
NE_D int ratio_walk(const DVolume& v, Tracker<BRICKMAJ>& trk, float& Tr, R& rng, Stats& st, int budget) {
	while (true) {
		if (budget-- <= 0) return TRACK_BUDGET;
		if (ratio_event<R, BRICKMAJ>(v, trk, Tr, rng, st) == TRACK_END) return TRACK_END;
	}
}
template <class R, bool BRICKMAJ>
NE_D int delta_walk(const DVolume& v, Tracker<BRICKMAJ>& trk, R& rng, Stats& st, int budget) {
	while (true) {
		if (budget-- <= 0) return TRACK_BUDGET;
		int e = delta_event<R, BRICKMAJ>(v, trk, rng, st);
		if (e != TRACK_MOVED) return e;
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
