// ne_tracking.cuh — GridMedia::Tr (ratio tracking, materials/GridMedia.cpp:45-69) and GridMedia::sample (delta
// tracking, :71-100) as RESUMABLE walks made of single EVENTS. Included by ne_device.cuh (needs density_at,
// brick_density, bsdf_sample).
//
// A Tracker walks the OCS ray segment [tStart, tFar]:
//   Tracker<false>  the reference's walk: t -= log(1 - xi) * invMaxDensity / sigma_bar against the single global majorant,
//                   draw for draw what GridMedia does (tape tests, NE_B200_RENDER_GLOBAL_MAJORANT).
//   Tracker<true>   production: the same exponential free flights measured in OPTICAL DEPTH against the piecewise-
//                   constant majorant of the 8^3 bricks the ray crosses. One exponential variate `tau` is drawn per
//                   candidate; crossing a brick only subtracts (length x sigma_bar x brick majorant) from it - no
//                   uniform, no logarithm, no density look-up - and empty bricks subtract nothing. The free-flight
//                   distribution is exactly that of delta tracking with a local majorant (unbiased).
// One event() = one brick crossing or one candidate. Because walks are sequences of independent events they can be
// stopped anywhere: the wavefront's tracking kernels hand a lane the next queued walk as soon as its own ends, and bound
// the events of one pass (`budget`); a stopped walk moves its origin to the point reached, and the next pass draws a
// fresh tau (memoryless).
#pragma once

namespace ne {

enum { TRACK_END = 0, TRACK_CANDIDATE = 1, TRACK_BUDGET = 2, TRACK_MOVED = 3 };

template <class W>
NE_D float exp_variate(W& wr) { return -logf(1 - wr.next()); }
// Production walks (per-brick majorants, their own PCG stream): the exponential variate through the hardware log2
// (2 instructions instead of ~12; relative error ~1e-7 of a free-flight length, far below the estimator's noise).
// The reference-order walk (TRACK_GLOBAL, tape tests) keeps logf above.
#ifndef NE_FAST_EXP_VARIATE
#define NE_FAST_EXP_VARIATE 1
#endif
template <class W>
NE_D float exp_variate_fast(W& wr) {
#if NE_FAST_EXP_VARIATE
	return -0.69314718056f * __log2f(1 - wr.next());
#else
	return -logf(1 - wr.next());
#endif
}

// MODE 0 (= false): the reference's global-majorant walk. 1 (= true): per-brick majorants. 2: per-brick majorants + empty-space
// skipping (TRACK_SKIP; the wavefront's tracking kernels pick it for sparse tables, see wavefront_render).
// 3: per-brick majorants read from a copy of the 2-byte table in SHARED memory (TRACK_BRICK_SM; the wavefront's tracking
// kernels stage it with a bulk async copy when it fits, and point DVolume::maj16 of their staged scene at it).
// 4: the same with empty-space skipping.
enum { TRACK_GLOBAL = 0, TRACK_BRICK = 1, TRACK_SKIP = 2, TRACK_BRICK_SM = 3, TRACK_SKIP_SM = 4 };
#ifndef NE_TRACK_FAST_INIT
#define NE_TRACK_FAST_INIT 1  // hardware reciprocals in the DDA set-up of production walks (C2: 23.85 -> 23.71 ms)
#endif
#define NE_TRACK_IS_SM(MODE) ((MODE) == TRACK_BRICK_SM || (MODE) == TRACK_SKIP_SM)
#define NE_TRACK_IS_SKIP(MODE) ((MODE) == TRACK_SKIP || (MODE) == TRACK_SKIP_SM)
template <int MODE>
struct Tracker;

template <>
struct Tracker<TRACK_GLOBAL> {
	Ray ray;  // OCS
	float t, tFar;
	float sig;     // sigma_bar per unit density = avg(extinction * densityMultiplier)
	float invMaj;  // GridMedia::invMaxDensity
	template <class W>
	NE_D void init(const DVolume& v, const DMaterial& m, Ray rayOCS, float tStart, float tEnd, W&, Stats&) {
		ray = rayOCS;
		t = tStart;
		tFar = tEnd;
		V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
		sig = avg(ext * m.density_mult);
		invMaj = v.inv_max_density;
	}
	// The three phases of one event (the wavefront kernels run them as separate warp-wide steps):
	//   wants_candidate  does the walk propose a collision point next (true) or does it have to move on (false)?
	//   move             cross into the next brick; TRACK_END when the segment is finished
	//   candidate_density  density at the proposed point (trk.t)
	template <class W>
	NE_D bool wants_candidate(W& wr) {
		t -= logf(1 - wr.next()) * invMaj / sig;  // GridMedia.cpp:56 / :82
		return t < tFar;
	}
	NE_D int move(Stats&) { return TRACK_END; }
	NE_D float candidate_density(const DVolume& v) { return density_at(v, ray, t); }
	template <class W>
	NE_D void after_candidate(W&) {}
};

// The per-brick majorant table a walk reads at every brick crossing (built by ne_bricks.cu, k_brick_table):
//   layout   (nbx+2) x (nby+2) x (nbz+2) entries of 2 bytes, x fastest: the bricks with a ONE-BRICK APRON all round.
//            Entry of brick (bx,by,bz) = table[(bz+1)*SZ + (by+1)*SY + (bx+1)], SY = nbx+2, SZ = SY*(nby+2).
//   entry    an IEEE half:  h > 0   the brick's majorant times 2^k, rounded UP (k per volume, DVolume::maj_scale = 2^-k),
//                                   so sigma_bar x majorant = (sig * maj_scale) * float(h): one conversion and one multiply
//                           h = -d  empty brick; d = Chebyshev distance in bricks (1..16) to the nearest brick with a record
//                                   or to the outside (TRACK_SKIP crosses the whole cube of empty bricks in one move)
//                           -inf    the apron: the walk has left the grid. Stepping out is therefore detected by the value
//                                   the step reads anyway - no per-axis bounds test in the loop.
// One brick crossing (move()) is ~30 instructions: 2 compares pick the axis (equality with the exit time, which IS one of
// nx/ny/nz), 3 predicated adds advance the brick and 3 the boundary times, 2 multiply-adds form the table index.
NE_D float half_bits_to_float(unsigned short h) { return __half2float(__ushort_as_half(h)); }

template <int MODE>
struct BrickTracker {
	const int2* __restrict__ cells;
	const float* __restrict__ pool;
	const unsigned short* __restrict__ tab;  // global-memory table (not TRACK_*_SM)
	uint32_t tabS;  // TRACK_*_SM: shared-window address of the table
	int SY, SZ;     // table strides (apron layout)
	int nbx, nby;   // brick grid (the un-aproned cells[] index of a candidate's record)
	V3 g0, gd;      // grid-space ray g(t) = g0 + t * gd
	float t, tFar;
	float tExit;    // where the ray leaves the current brick (clipped to tFar)
	float sig;      // sigma_bar per unit density per unit t
	float sigK;     // sig * 2^-k: times the table's half gives sigma_bar x majorant
	float sigMaj;   // sigma_bar x majorant of the current brick: optical depth per unit t (0 in an empty brick)
	float invMaj;   // 1 / majorant, set when a candidate is proposed
	float tau;      // optical depth left before the next candidate
	float cap;      // scratch of wants_candidate(): optical depth of the rest of the current brick
	bool out;       // the last step left the grid
	// brick DDA: current brick, step per axis (+-1), ray parameter of the next boundary crossing per axis and its period
	int bx, by, bz, sx, sy, sz;
	float nx, ny, nz, dx, dy, dz;

	template <class W>
	NE_D void init(const DVolume& v, const DMaterial& m, Ray rayOCS, float tStart, float tEnd, W& wr, Stats& st) {
		cells = v.cells;
		pool = v.pool;
		tab = v.maj16;
		if (NE_TRACK_IS_SM(MODE)) tabS = uint32_t(__cvta_generic_to_shared(v.maj16));
		nbx = v.bx; nby = v.by;
		SY = v.bx + 2;
		SZ = SY * (v.by + 2);
		V3 res(float(v.W), float(v.H), float(v.D));
		g0 = (rayOCS.o + V3(0.5f)) * res;
		gd = rayOCS.d * res;
		t = tStart;
		tFar = tEnd;
		V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
#if NE_TRACK_FAST_INIT
		{
			V3 e = ext * m.density_mult;
			sig = (e.x + e.y + e.z) * (1.0f / 3.0f);
		}
#else
		sig = avg(ext * m.density_mult);
#endif
		sigK = sig * v.maj_scale;
		// DDA set-up at the segment's first point
		V3 g = point(tStart);
		bx = min(max(int(floorf(g.x * 0.125f)), 0), v.bx - 1);
		by = min(max(int(floorf(g.y * 0.125f)), 0), v.by - 1);
		bz = min(max(int(floorf(g.z * 0.125f)), 0), v.bz - 1);
#if NE_TRACK_FAST_INIT
		// (production walks only: where a DDA boundary falls by an ulp is immaterial to the estimate; the reference-order walk
		// is Tracker<TRACK_GLOBAL>)
		float ix = __fdividef(1.0f, gd.x), iy = __fdividef(1.0f, gd.y), iz = __fdividef(1.0f, gd.z);  // +-inf for an axis-parallel ray
#else
		float ix = 1.0f / gd.x, iy = 1.0f / gd.y, iz = 1.0f / gd.z;  // +-inf for an axis-parallel ray
#endif
		sx = gd.x > 0 ? 1 : -1; sy = gd.y > 0 ? 1 : -1; sz = gd.z > 0 ? 1 : -1;
		dx = gd.x != 0 ? 8.0f * fabsf(ix) : INFINITY;
		dy = gd.y != 0 ? 8.0f * fabsf(iy) : INFINITY;
		dz = gd.z != 0 ? 8.0f * fabsf(iz) : INFINITY;
		nx = gd.x != 0 ? (float((gd.x > 0 ? bx + 1 : bx) << 3) - g.x) * ix + tStart : INFINITY;
		ny = gd.y != 0 ? (float((gd.y > 0 ? by + 1 : by) << 3) - g.y) * iy + tStart : INFINITY;
		nz = gd.z != 0 ? (float((gd.z > 0 ? bz + 1 : bz) << 3) - g.z) * iz + tStart : INFINITY;
		enter_brick(st);
		tau = exp_variate_fast(wr);
	}
	NE_D V3 point(float tt) const { return V3(fmaf(gd.x, tt, g0.x), fmaf(gd.y, tt, g0.y), fmaf(gd.z, tt, g0.z)); }
	NE_D float exit_t() const { return fminf(nx, fminf(ny, nz)); }
	// Reads the brick's table entry (see the layout above): two bytes from shared memory or L1.
	NE_D void enter_brick(Stats& st) {
		st.brick_visits++;
		const int idx = (bz + 1) * SZ + (by + 1) * SY + (bx + 1);
		unsigned short h;
		if (NE_TRACK_IS_SM(MODE)) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(tabS + 2u * uint32_t(idx)));
		else h = __ldg(tab + idx);
		const float f = half_bits_to_float(h);
		out = f == -INFINITY;
		if (NE_TRACK_IS_SKIP(MODE) && f < -1.0f && !out) jump(int(-f) - 1);
		sigMaj = fmaxf(sigK * f, 0.0f);  // empty bricks (negative entries) have nothing to collide with
		tExit = fminf(exit_t(), tFar);
	}
	template <class W>
	NE_D bool wants_candidate(W&) {
		cap = (tExit - t) * sigMaj;  // optical depth of the rest of this brick
		return tau < cap;
	}
	NE_D int move(Stats& st) {
		tau -= cap;
		t = tExit;
		if (tExit >= tFar) return TRACK_END;
		// cross the nearest boundary: tExit < tFar here, so it IS one of nx / ny / nz
		const bool cx = nx == tExit;
		const bool cy = !cx && ny == tExit;
		const bool cz = !cx && !cy;
		if (cx) { bx += sx; nx += dx; }
		if (cy) { by += sy; ny += dy; }
		if (cz) { bz += sz; nz += dz; }
		enter_brick(st);
		return out ? TRACK_END : TRACK_MOVED;
	}
	// Empty-space skip: every brick within Chebyshev distance r of the current one is empty (and inside the grid). Move the
	// DDA, in one go, to the LAST brick the ray visits inside that cube, so that exit_t() is where it leaves the cube and the
	// next move() crosses the cube's face. Per axis the ray crosses at most r boundaries before that moment: the exit axis
	// exactly r (its (r+1)-th crossing IS the exit), the others as many as lie before the exit time.
	NE_D void jump(int r) {
		const float fr = float(r);
		const float tx = fmaf(fr, dx, nx), ty = fmaf(fr, dy, ny), tz = fmaf(fr, dz, nz);  // inf for an axis-parallel ray
		const float tc = fminf(tx, fminf(ty, tz));
		int kx = nx <= tc ? min(r, int(__fdividef(tc - nx, dx)) + 1) : 0;
		int ky = ny <= tc ? min(r, int(__fdividef(tc - ny, dy)) + 1) : 0;
		int kz = nz <= tc ? min(r, int(__fdividef(tc - nz, dz)) + 1) : 0;
		bx += sx * kx;
		by += sy * ky;
		bz += sz * kz;
		nx = kx ? fmaf(float(kx), dx, nx) : nx;
		ny = ky ? fmaf(float(ky), dy, ny) : ny;
		nz = kz ? fmaf(float(kz), dz, nz) : nz;
	}
	// Only a candidate needs the brick's record: its slot is looked up here (sigMaj > 0, so the brick has one).
	NE_D float candidate_density(const DVolume&) {
		const float r = __fdividef(1.0f, sigMaj);
		invMaj = sig * r;
		t = fmaf(tau, r, t);
		const int slot = __ldg(&cells[(bz * nby + by) * nbx + bx].x);
		return brick_density(pool, slot, point(t), bx, by, bz);
	}
	template <class W>
	NE_D void after_candidate(W& wr) { tau = exp_variate_fast(wr); }
};
template <>
struct Tracker<TRACK_BRICK> : BrickTracker<TRACK_BRICK> {};
template <>
struct Tracker<TRACK_SKIP> : BrickTracker<TRACK_SKIP> {};
template <>
struct Tracker<TRACK_BRICK_SM> : BrickTracker<TRACK_BRICK_SM> {};
template <>
struct Tracker<TRACK_SKIP_SM> : BrickTracker<TRACK_SKIP_SM> {};

// One event: TRACK_CANDIDATE (`dens` = density at the proposed collision point trk.t), TRACK_MOVED, or TRACK_END.
template <class W, int BRICKMAJ>
NE_D int track_advance(const DVolume& v, Tracker<BRICKMAJ>& trk, W& wr, float& dens, Stats& st) {
	if (trk.wants_candidate(wr)) {
		dens = trk.candidate_density(v);
		return TRACK_CANDIDATE;
	}
	return trk.move(st);
}

#define NE_NO_BUDGET 0x7fffffff

// What a candidate does to a ratio-tracking walk, with pbrt's Russian roulette (GridMedia.cpp:58-66). TRACK_END =
// killed (Tr = 0), TRACK_MOVED = keep going.
template <class W, int BRICKMAJ>
NE_D int ratio_candidate(Tracker<BRICKMAJ>& trk, float density, float& Tr, W& wr, Stats& st) {
	st.ratio_steps++;
	Tr *= 1 - fmaxf(0.0f, density * trk.invMaj);
	const float rrThreshold = .1f;
	if (Tr < rrThreshold) {
		float q = fmaxf(0.05f, 1.0f - Tr);
		if (wr.next() < q) { Tr = 0.0f; return TRACK_END; }
		Tr /= 1 - q;
	}
	trk.after_candidate(wr);
	return TRACK_MOVED;
}
// What a candidate does to a delta-tracking walk (GridMedia.cpp:84-95). TRACK_CANDIDATE = REAL collision at trk.t.
template <class W, int BRICKMAJ>
NE_D int delta_candidate(Tracker<BRICKMAJ>& trk, float density, W& wr, Stats& st) {
	st.delta_steps++;
	float ra = wr.next();
	if (density * trk.invMaj > ra) return TRACK_CANDIDATE;
	trk.after_candidate(wr);
	return TRACK_MOVED;
}
// One ratio-tracking event. Returns TRACK_END when the walk is over (Tr final, possibly 0 = killed), else TRACK_MOVED.
template <class W, int BRICKMAJ>
NE_D int ratio_event(const DVolume& v, Tracker<BRICKMAJ>& trk, float& Tr, W& wr, Stats& st) {
	float density;
	int e = track_advance(v, trk, wr, density, st);
	if (e != TRACK_CANDIDATE) return e;
	return ratio_candidate(trk, density, Tr, wr, st);
}
// One delta-tracking event. TRACK_CANDIDATE = REAL collision at trk.t, TRACK_END = escaped, TRACK_MOVED = keep going.
template <class W, int BRICKMAJ>
NE_D int delta_event(const DVolume& v, Tracker<BRICKMAJ>& trk, W& wr, Stats& st) {
	float density;
	int e = track_advance(v, trk, wr, density, st);
	if (e != TRACK_CANDIDATE) return e;
	return delta_candidate(trk, density, wr, st);
}

template <class W, int BRICKMAJ>
NE_D int ratio_walk(const DVolume& v, Tracker<BRICKMAJ>& trk, float& Tr, W& wr, Stats& st, int budget) {
	while (true) {
		if (budget-- <= 0) return TRACK_BUDGET;
		if (ratio_event<W, BRICKMAJ>(v, trk, Tr, wr, st) == TRACK_END) return TRACK_END;
	}
}
template <class W, int BRICKMAJ>
NE_D int delta_walk(const DVolume& v, Tracker<BRICKMAJ>& trk, W& wr, Stats& st, int budget) {
	while (true) {
		if (budget-- <= 0) return TRACK_BUDGET;
		int e = delta_event<W, BRICKMAJ>(v, trk, wr, st);
		if (e != TRACK_MOVED) return e;
	}
}

// GridMedia::Tr. rayW: WCS ray; tNear/tFar from the hit record. Returns the scalar transmittance.
template <class R, bool BRICKMAJ>
NE_D float grid_tr(const DInstance& in, const DMaterial& m, const DVolume& v, Ray rayW, float tNear, float tFar, R& rng, Stats& st) {
	Ray ray = transform_ray(rayW, in.Mi);
	ray.o = ray.at(tNear);
	typename WalkRngOf<(BRICKMAJ != 0), R>::type wr;
	wr.start(rng);
	Tracker<BRICKMAJ> trk;
	trk.init(v, m, ray, 0.0f, tFar - tNear, wr, st);
	float Tr = 1;
	ratio_walk(v, trk, Tr, wr, st, NE_NO_BUDGET);
	return Tr;
}

// The scattering vertex of GridMedia::sample (:90-95) at parameter t of the OCS ray: phase sample about +Y of the
// OCS (Q19: ignores the incoming direction), mapped to WCS by M (the direction keeps the instance scale).
template <bool FAST = false, class R>
NE_D Ray grid_scatter(const DScene& s, const DInstance& in, const DMaterial& m, const Ray& rayOCS, float t, const Hit& isect, R& rng, Stats& st) {
	st.scatter_events++;
	Ray so;
	so.o = rayOCS.at(t);
	so.d = bsdf_sample<1, FAST>(s, m, rayOCS.d, V3(0.0f, 1.0f, 0.0f), isect, rng);
	return transform_ray(so, in.M);
}

// GridMedia::sample. In Li `incoming` already has its origin at the segment start and tNear = 0.
// Returns the value the reference returns: sigma_s/sigma_t on a collision, exactly (1,1,1) on escape (Q1, Q1b).
template <class R, bool BRICKMAJ>
NE_D V3 grid_sample(const DScene& s, const DInstance& in, const DMaterial& m, const DVolume& v, Ray incomingW, float tNear, float tFar,
                    const Hit& isect, Ray& scattered, R& rng, Stats& st) {
	scattered = incomingW;
	Ray ray = transform_ray(incomingW, in.Mi);
	typename WalkRngOf<(BRICKMAJ != 0), R>::type wr;
	wr.start(rng);
	Tracker<BRICKMAJ> trk;
	trk.init(v, m, ray, tNear, tFar, wr, st);  // GridMedia.cpp:79 starts at tNear; Li always passes 0
	if (delta_walk(v, trk, wr, st, NE_NO_BUDGET) != TRACK_CANDIDATE) return V3(1.0f);
	scattered = grid_scatter(s, in, m, ray, trk.t, isect, rng, st);
	V3 sc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
	V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + sc;
	return sc / ext;
}

}  // namespace ne
