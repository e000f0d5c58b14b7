// ne_wavefront.cu — the production renderer: OfflineEngine::renderTile's pixel x sample loops
// (core/OfflineEngine.cpp:61-71) over the whole frame as a WAVEFRONT of path records.
//
//   pool      N path slots (one 128-byte record each: 64 B path state + 48 B hit) that are refilled with new camera
//             samples as paths terminate, so the wavefront stays full until the work runs out. 64 Mi slots by default
//             (21 GB with the request arrays: HBM is plentiful, and a wide wavefront means few, long launches)
//   queues    arrays of slot indices: {generate, extend} -> {volume, scatter (homogeneous media), surface};
//             volume -> {scatter, volume (walk not finished), next extend}; free slots. Pushes are warp-aggregated
//             (one atomicAdd per warp per queue; the persistent tracking kernels batch theirs per phase)
//   kernels   plan / commit (1 thread: queue bookkeeping)
//             · generate (camera ray + the path's first Scene::intersectScene: paths that end at once never enter the pool)
//             · extend (Scene::intersectScene fold, BVH, for continuing paths)
//             · track (delta tracking: persistent warps in finish+refill / move / candidate phases, bounded events per pass)
//             · scatter (phase function + next-event setup) · surface (GGX shading + next-event setup)
//             · shadow (visibilityTr requests) · trfind (intersectTr: walk through surfaces to the first medium)
//             · tr (ratio tracking through that medium, persistent warps like track)
//             · trace<Extend|Shadow|TrFind job> (scenes with meshes: the three ray-casting stages as persistent warps over
//               a resumable intersectScene, so a warp is not held by its longest BVH walk)
//             variants chosen per scene by the host: scatter<FUSE> traces its own continuation ray in mesh-free scenes;
//             track/tr<TRACK_SKIP> cross cubes of empty bricks in one move where the brick table is sparse
//   output    fp32 atomicAdd splats into the context's linear accumulation buffer
//
// The two tracking kernels stop a walk after `budget` events (brick crossings + density look-ups) and queue the
// remainder for the next pass: exponential free flights are memoryless, so the estimate is unchanged, and a warp is
// never held hostage by its longest walk. Every kernel runs over a device-side count with a fixed grid of (SM count x
// resident blocks): no host round trip sizes a launch; the host only polls a mapped "done" word every few iterations.
// Every block stages the scene's instance / material / volume tables in shared memory first (stage_scene).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ne_ctx.h"
#include "ne_integrator.cuh"

using namespace ne;

namespace {

struct WfCounts {
	uint32_t extend, next, vol, volNext, volHead, scat, surf, freeN, shadow, tr, trNext, trHead, trNew0, gen, genTaken, done, extHead, shHead, trfHead, pad_;
	unsigned long long workNext, workTotal;
};

struct __align__(128) PathSlot {
	float4 pA, pB, pC;
	uint4 pD;
	float4 hA, hB, hC;
	float4 spare;
};

struct WfBuf {
	// One 128-byte record per path slot (path state + hit), so that a slot reached through a queue index - a random
	// address - costs exactly one L2 line, every byte of it used:
	//   path  pA=(o.xyz,d.x) pB=(d.yz,T.xy) pC=(T.z,pixel,sample,dim) pD=(bounce|guard<<8, nee, collision t, -)
	//   hit   hA=(p.xyz,tNear) hB=(n.xyz,tFar) hC=(u,v,inst,prim)
	PathSlot* rec;
	// shadow request: A=(o.xyz,C.x) B=(C.yz,w.xy) C=(w.z,pixel)
	float4 *sA, *sB;
	float2* sC;
	// transmittance request (current / next pass): A=(o.xyz,d.x) B=(d.yz,w.xy) C=(w.z,pixel,sample,stream)
	// D=(Tr so far, remaining tFar, medium instance or -1 = not found yet, dim)
	float4 *tA, *tB, *tC, *tD;
	float4 *uA, *uB, *uC, *uD;
	uint32_t *qExtend, *qNext, *qVol, *qVolNext, *qScat, *qSurf, *qFree;
	WfCounts* c;
};

struct WfParams {
	DScene s;
	DCamera cam;
	float* accum;
	int W, H, sppBegin, bounces, budget, refill, moves, walkBudget, walkRefill;
	uint64_t seed;
	DCounters* counters;
};

__device__ __forceinline__ uint32_t warp_push(uint32_t* counter) {
	unsigned m = __activemask();
	unsigned lane = threadIdx.x & 31;
	int leader = __ffs(m) - 1;
	uint32_t base = 0;
	if (int(lane) == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
	base = __shfl_sync(m, base, leader);
	return base + __popc(m & ((1u << lane) - 1));
}

__device__ __forceinline__ void flush_stats_wf(const Stats& st, DCounters* c) {
	unsigned m = __activemask();
	unsigned lane = threadIdx.x & 31;
	unsigned leader = __ffs(m) - 1;
#define NE_FLUSH(field)                                                        \
	{                                                                          \
		unsigned v = __reduce_add_sync(m, (unsigned)(st.field));               \
		if (lane == leader && v) atomicAdd(&c->field, (unsigned long long)v);  \
	}
	NE_FLUSH(extend_rays)
	NE_FLUSH(shadow_rays)
	NE_FLUSH(delta_steps)
	NE_FLUSH(ratio_steps)
	NE_FLUSH(brick_visits)
	NE_FLUSH(bvh_nodes)
	NE_FLUSH(tri_tests)
	NE_FLUSH(prim_tests)
	NE_FLUSH(scatter_events)
	NE_FLUSH(surface_events)
#undef NE_FLUSH
}

__device__ __forceinline__ void splat(float* accum, uint32_t pixel, V3 v) {
	if (v.x != 0) atomicAdd(accum + 3 * size_t(pixel), v.x);
	if (v.y != 0) atomicAdd(accum + 3 * size_t(pixel) + 1, v.y);
	if (v.z != 0) atomicAdd(accum + 3 * size_t(pixel) + 2, v.z);
}

struct PathRec {
	PathState ps;
	uint32_t pixel, sample, dim;
	float tHit;  // collision parameter handed from track to scatter
};
__device__ __forceinline__ PathRec load_path(const WfBuf& b, uint32_t slot) {
	float4 A = b.rec[slot].pA, B = b.rec[slot].pB, C = b.rec[slot].pC;
	uint4 D = b.rec[slot].pD;
	PathRec r;
	r.ps.ray.o = V3(A.x, A.y, A.z);
	r.ps.ray.d = V3(A.w, B.x, B.y);
	r.ps.T = V3(B.z, B.w, C.x);
	r.pixel = __float_as_uint(C.y);
	r.sample = __float_as_uint(C.z);
	r.dim = __float_as_uint(C.w);
	r.ps.bounce = int(D.x & 0xff);
	r.ps.guard = int(D.x >> 8);
	r.ps.nee = D.y;
	r.tHit = __uint_as_float(D.z);
	return r;
}
__device__ __forceinline__ void store_path(const WfBuf& b, uint32_t slot, const PathRec& r) {
	b.rec[slot].pA = make_float4(r.ps.ray.o.x, r.ps.ray.o.y, r.ps.ray.o.z, r.ps.ray.d.x);
	b.rec[slot].pB = make_float4(r.ps.ray.d.y, r.ps.ray.d.z, r.ps.T.x, r.ps.T.y);
	b.rec[slot].pC = make_float4(r.ps.T.z, __uint_as_float(r.pixel), __uint_as_float(r.sample), __uint_as_float(r.dim));
	b.rec[slot].pD = make_uint4(uint32_t(r.ps.bounce) | (uint32_t(r.ps.guard) << 8), r.ps.nee, __float_as_uint(r.tHit), 0u);
}
__device__ __forceinline__ Hit load_hit(const WfBuf& b, uint32_t slot) {
	float4 A = b.rec[slot].hA, B = b.rec[slot].hB, C = b.rec[slot].hC;
	Hit h;
	h.p = V3(A.x, A.y, A.z);
	h.tNear = A.w;
	h.n = V3(B.x, B.y, B.z);
	h.tFar = B.w;
	h.u = C.x;
	h.v = C.y;
	h.inst = __float_as_int(C.z);
	h.prim = __float_as_int(C.w);
	return h;
}
__device__ __forceinline__ void store_hit(const WfBuf& b, uint32_t slot, const Hit& h) {
	b.rec[slot].hA = make_float4(h.p.x, h.p.y, h.p.z, h.tNear);
	b.rec[slot].hB = make_float4(h.n.x, h.n.y, h.n.z, h.tFar);
	b.rec[slot].hC = make_float4(h.u, h.v, __int_as_float(h.inst), __int_as_float(h.prim));
}

// The instance / material / volume tables are a few hundred bytes that every ray of every kernel reads through two or
// three DEPENDENT loads (instance -> material -> volume). Each block stages them in shared memory once, so those
// look-ups cost a shared-memory access instead of an L1/L2 round trip (36 % of k_wf_extend's stall samples before).
// Scenes with more entries than fit keep reading them from global memory.
#define NE_CACHE_INST 24
#define NE_CACHE_MAT 24
#define NE_CACHE_VOL 8
struct SceneCache {
	DInstance inst[NE_CACHE_INST];
	DMaterial mat[NE_CACHE_MAT];
	DVolume vol[NE_CACHE_VOL];
};
__device__ __forceinline__ DScene stage_scene(const DScene& g, SceneCache& sh) {
	DScene s = g;
	if (g.n_inst <= NE_CACHE_INST && g.n_mat <= NE_CACHE_MAT && g.n_vol <= NE_CACHE_VOL) {
		const uint32_t* src;
		uint32_t* dst;
		src = reinterpret_cast<const uint32_t*>(g.inst); dst = reinterpret_cast<uint32_t*>(sh.inst);
		for (uint32_t k = threadIdx.x; k < g.n_inst * (sizeof(DInstance) / 4); k += blockDim.x) dst[k] = src[k];
		src = reinterpret_cast<const uint32_t*>(g.mat); dst = reinterpret_cast<uint32_t*>(sh.mat);
		for (uint32_t k = threadIdx.x; k < g.n_mat * (sizeof(DMaterial) / 4); k += blockDim.x) dst[k] = src[k];
		src = reinterpret_cast<const uint32_t*>(g.vol); dst = reinterpret_cast<uint32_t*>(sh.vol);
		for (uint32_t k = threadIdx.x; k < g.n_vol * (sizeof(DVolume) / 4); k += blockDim.x) dst[k] = src[k];
		__syncthreads();
		s.inst = sh.inst;
		s.mat = sh.mat;
		s.vol = sh.vol;
	}
	return s;
}
#define NE_STAGE_SCENE()               \
	__shared__ SceneCache sceneCache_; \
	const DScene S = stage_scene(P.s, sceneCache_)

// Turns the two next-event queries of estimateDirect into requests; emission and request weights go straight to
// the accumulation buffer / request arrays.
struct QueueSink {
	V3 scale;
	float sel_pdf;
	const WfBuf* b;
	float* accum;
	uint32_t pixel, sample;
	__device__ __forceinline__ void begin() {}
	__device__ __forceinline__ void emit(V3 v) { splat(accum, pixel, v); }
	__device__ __forceinline__ V3 end(float) { return V3(0.0f); }
	__device__ __forceinline__ void light_term(const DScene&, V3 p, V3 C, V3 f, V3 Li, float weight, float pdf, PhiloxRng&, Stats&) {
		V3 w = scale * ((f * Li * weight / pdf) / sel_pdf);
		if (is_black(w)) return;
		uint32_t i = warp_push(&b->c->shadow);
		b->sA[i] = make_float4(p.x, p.y, p.z, C.x);
		b->sB[i] = make_float4(C.y, C.z, w.x, w.y);
		b->sC[i] = make_float2(w.z, __uint_as_float(pixel));
	}
	__device__ __forceinline__ void bsdf_term(const DScene& s, Ray ray, V3 f, V3 Li, float weight, float pdf, PhiloxRng&, uint32_t stream, Stats&) {
		if (!s.has_medium) return;  // intersectTr can only succeed through a medium (Q12)
		V3 w = scale * ((f * Li * weight / pdf) / sel_pdf);
		if (is_black(w)) return;
		uint32_t i = warp_push(&b->c->tr);
		b->tA[i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
		b->tB[i] = make_float4(ray.d.y, ray.d.z, w.x, w.y);
		b->tC[i] = make_float4(w.z, __uint_as_float(pixel), __uint_as_float(sample), __uint_as_float(stream));
		b->tD[i] = make_float4(1.0f, 0.0f, __int_as_float(-1), __uint_as_float(0u));
	}
};

// ---------------------------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_wf_init(WfBuf b, uint32_t nSlots, unsigned long long workTotal) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nSlots) b.qFree[i] = nSlots - 1 - i;  // slot 0 is handed out first
	if (i == 0) {
		WfCounts c;
		memset(&c, 0, sizeof(c));
		c.freeN = nSlots;
		c.workTotal = workTotal;
		*b.c = c;
	}
}

// One thread: retire the finished iteration (next -> extend, unfinished walks -> vol / tr, clear stage queues) and
// plan the refill. The host alternates which buffer of each ping-pong pair is "current".
__global__ void k_wf_plan(WfBuf b, volatile uint32_t* hostDone) {
	WfCounts& c = *b.c;
	c.extend = c.next;
	c.next = 0;
	c.vol = c.volNext;
	c.volNext = 0;
	c.tr = c.trNext;
	c.trNew0 = c.trNext;  // requests pushed from here on are new: k_wf_trfind locates their medium
	c.trNext = 0;
	c.scat = c.surf = c.shadow = 0;
	c.volHead = c.trHead = c.extHead = c.shHead = c.trfHead = 0;
	unsigned long long remaining = c.workTotal - c.workNext;
	uint32_t gen = uint32_t(remaining < c.freeN ? remaining : c.freeN);
	c.gen = gen;
	c.freeN -= gen;
	c.done = (c.extend == 0 && gen == 0 && c.vol == 0 && c.tr == 0) ? 1u : 0u;
	if (hostDone) *hostDone = c.done;
}

// OfflineEngine.cpp:64-67 + Li's first intersectScene (:187-193, :244-260) fused: sample jitter,
// Camera::getRayPassingThrough, trace and classify `gen` new camera paths. A camera path that ends right there (it
// misses everything, or sees an emitter) never touches the pool - no slot, no record, no queue entry; only survivors
// take a slot (from the `gen` the plan set aside; commit returns the rest) and are written ONCE, path and hit
// together, straight into the volume / surface queue. In the C2 frame 5 of 6 camera paths miss the medium's box.
template <int MINB>  // resident blocks per SM: 3 without meshes, 4 with (the BVH walk is latency-bound: warps in flight count)
__global__ void __launch_bounds__(256, MINB) k_wf_generate(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t gen = b.c->gen;
	const uint32_t freeN = b.c->freeN;
	const unsigned long long workBase = b.c->workNext;
	const uint32_t npix = uint32_t(P.W) * uint32_t(P.H);
	const bool tiled = (P.W % 8 == 0) && (P.H % 4 == 0);  // else the frame is walked row by row (any bijection will do: Philox is keyed by pixel)
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < gen; i += gridDim.x * blockDim.x) {
		unsigned long long w = workBase + i;
		uint32_t pixel = uint32_t(w % npix);
		if (tiled) {  // a warp's 32 consecutive work items cover an 8x4 pixel tile, not a 32x1 strip: coherent camera rays
			const uint32_t t = pixel >> 5, l = pixel & 31u, tilesX = uint32_t(P.W) >> 3;
			pixel = ((t / tilesX) * 4u + (l >> 3)) * uint32_t(P.W) + (t % tilesX) * 8u + (l & 7u);
		}
		uint32_t sample = uint32_t(P.sppBegin) + uint32_t(w / npix);
		int x = int(pixel % uint32_t(P.W)), y = int(pixel / uint32_t(P.W));
		PhiloxRng rng;
		rng.init(P.seed, pixel, sample);
		float u = float(float(x) + rng.next()) / float(P.W);
		float v = float(float(y) + rng.next()) / float(P.H);
		PathRec r;
		r.ps.ray = camera_ray(P.cam, u, v, rng);
		r.ps.T = V3(1.0f);
		r.ps.bounce = 0;
		r.ps.guard = 0;
		r.ps.nee = 0;
		r.pixel = pixel;
		r.sample = sample;
		r.dim = rng.dim;
		r.tHit = 0;
		Hit h;
		st.extend_rays++;
		bool did = intersect_scene(S, r.ps.ray, h, float(NE_EPSILON12), INFINITY, st);
		QueueSink sink;
		sink.accum = P.accum;
		sink.pixel = pixel;
		int kind = classify_hit(S, did, h, r.ps, sink);
		if (kind == HIT_TERMINATE) continue;
		uint32_t slot = b.qFree[freeN + gen - 1 - warp_push(&b.c->genTaken)];
		store_path(b, slot, r);
		store_hit(b, slot, h);
		if (kind == HIT_VOLUME) {
			if (S.mat[S.inst[h.inst].material].volume >= 0) b.qVol[warp_push(&b.c->vol)] = slot;
			else b.qScat[warp_push(&b.c->scat)] = slot;
		} else b.qSurf[warp_push(&b.c->surf)] = slot;
	}
	flush_stats_wf(st, P.counters);
}
// One thread: publish the refill (after generate has read the old counts); unused reserved slots go back to the free list.
__global__ void k_wf_commit(WfBuf b, DCounters* counters) {
	WfCounts& c = *b.c;
	c.freeN += c.gen - c.genTaken;
	c.workNext += c.gen;
	atomicAdd(&counters->paths, (unsigned long long)c.gen);
	c.gen = 0;
	c.genTaken = 0;
}

// Scene::intersectScene for every path of the extend queue + classify (Li :187-193, :244-260).
__global__ void __launch_bounds__(256) k_wf_extend(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t n = b.c->extend;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qExtend[i];
		PathRec r = load_path(b, slot);
		Hit h;
		st.extend_rays++;
		bool did = intersect_scene(S, r.ps.ray, h, float(NE_EPSILON12), INFINITY, st);
		QueueSink sink;
		sink.accum = P.accum;
		sink.pixel = r.pixel;
		int kind = classify_hit(S, did, h, r.ps, sink);
		if (kind == HIT_TERMINATE) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			store_hit(b, slot, h);
			if (kind == HIT_VOLUME) {
				// a grid medium is walked by k_wf_track; a HomogeneousMedia needs no walk and goes straight to k_wf_scatter
				if (S.mat[S.inst[h.inst].material].volume >= 0) b.qVol[warp_push(&b.c->vol)] = slot;
				else b.qScat[warp_push(&b.c->scat)] = slot;
			} else b.qSurf[warp_push(&b.c->surf)] = slot;
		}
	}
	flush_stats_wf(st, P.counters);
}

#ifndef NE_TRACK_EARLY_BREAK
#define NE_TRACK_EARLY_BREAK 0  // leave the move loop as soon as no lane of the warp is moving (a vote per crossing: measured slower)
#endif
#ifndef NE_TRACK_THREADS
#define NE_TRACK_THREADS 256
#endif
#ifndef NE_TRACK_BLOCKS
#define NE_TRACK_BLOCKS 4  // resident blocks per SM the tracking kernels are compiled for (64 registers per thread)
#endif

// Lane states of the persistent tracking kernels.
enum { L_IDLE = 0, L_MOVING = 1, L_CAND = 2, L_FIN_HIT = 3, L_FIN_BUDGET = 4, L_FIN_END = 5 };

// Reserve `popc(mask)` entries of a queue for the lanes in `mask`: one atomicAdd per warp. All 32 lanes call it.
// Returns this lane's entry (meaningful for lanes in the mask). The atomics of consecutive calls are independent,
// so their round trips to L2 overlap; the shuffle that needs the result comes after all of them have been issued.
struct WarpReserve {
	uint32_t base;
	unsigned mask;
	__device__ __forceinline__ void issue(uint32_t* counter, bool mine) {
		mask = __ballot_sync(0xffffffffu, mine);
		base = 0;
		if ((threadIdx.x & 31) == 0 && mask) base = atomicAdd(counter, (uint32_t)__popc(mask));
	}
	__device__ __forceinline__ uint32_t get() const {
		uint32_t b0 = __shfl_sync(0xffffffffu, base, 0);
		return b0 + __popc(mask & ((1u << (threadIdx.x & 31)) - 1));
	}
};

// Delta tracking (GridMedia::sample's loop) for every path of the volume queue, by PERSISTENT warps that run three
// warp-wide phases in turn so that lanes doing the same kind of work run together:
//   finish + refill  (when P.refill lanes are not walking) finished walks write back the fields they changed and are
//                    queued for the next stage; idle lanes take the next queued walks. All queue atomics of the phase
//                    (three pushes and the fetch) are issued back to back, one L2 round trip for the lot.
//   move             up to P.moves brick crossings per walking lane (one 2-byte majorant load each, no density)
//   candidate        every lane that proposed a collision point looks the density up (eight loads from one brick
//                    record) and accepts or rejects it
// A walk ends on a real collision, on leaving the medium, or after P.budget events (it then continues from the point
// reached in the next pass), so a warp is never left with one lane grinding through a long walk while 31 idle.
template <int BRICKMAJ>  // TRACK_GLOBAL / TRACK_BRICK / TRACK_SKIP (ne_tracking.cuh)
__global__ void __launch_bounds__(NE_TRACK_THREADS, NE_TRACK_BLOCKS) k_wf_track(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	typedef typename WalkRngOf<(BRICKMAJ != 0), PhiloxRng>::type WalkRng;
	const uint32_t n = b.c->vol;
	Stats st;
	st.clear();
	int state = L_IDLE;
	bool exhausted = false;
	uint32_t slot = 0, bg = 0;
	Ray ray;       // WCS, origin at the start of the segment
	float tFar = 0;
	PhiloxRng rng;
	WalkRng wr;
	Tracker<BRICKMAJ> trk;
	const DVolume* vol = nullptr;
	int budget = 0;
	while (true) {
		unsigned walking = __ballot_sync(0xffffffffu, state == L_MOVING || state == L_CAND);
		if (walking == 0 || (!exhausted && 32 - __popc(walking) >= P.refill)) {
			// ---- finish + refill
			const bool fin = state >= L_FIN_HIT;
			const bool esc = state == L_FIN_END;
			const uint32_t bg2 = bg + 256u;  // volume_escape (Li :209-213, Q1): the guard lives in bits 8..31
			const bool dead = esc && (bg2 >> 8) > NE_MAX_NULL_SEGMENTS;
			WarpReserve rScat, rVolNext, rNext, rFree, rFetch;
			rScat.issue(&b.c->scat, state == L_FIN_HIT);
			rVolNext.issue(&b.c->volNext, state == L_FIN_BUDGET);
			rNext.issue(&b.c->next, esc && !dead);
			rFree.issue(&b.c->freeN, dead);
			rFetch.issue(&b.c->volHead, !exhausted && (state == L_IDLE || fin));
			// the shuffles are warp-wide: resolve every reservation before the lanes part ways
			const uint32_t iScat = rScat.get(), iVolNext = rVolNext.get(), iNext = rNext.get(), iFree = rFree.get(), iFetch = rFetch.get();
			if (fin) {
				if (state == L_FIN_HIT) {
					b.rec[slot].pA = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
					b.rec[slot].pD.z = __float_as_uint(trk.t);
					b.rec[slot].hB.w = tFar;
					b.qScat[iScat] = slot;
				} else if (state == L_FIN_BUDGET) {
					V3 o = ray.at(trk.t);
					b.rec[slot].pA = make_float4(o.x, o.y, o.z, ray.d.x);
					b.rec[slot].hB.w = tFar - trk.t;
					b.qVolNext[iVolNext] = slot;
				} else if (dead) {
					b.qFree[iFree] = slot;
				} else {
					V3 o = ray.at(tFar + 0.01f);  // step past the far side, same bounce
					b.rec[slot].pA = make_float4(o.x, o.y, o.z, ray.d.x);
					b.rec[slot].pD.x = bg2;
					b.qNext[iNext] = slot;
				}
				b.rec[slot].pC.w = __uint_as_float(rng.dim);
				b.rec[slot].hA.w = 0.0f;
				state = L_IDLE;
			}
			if (!exhausted) {
				const uint32_t i = iFetch;
				if (state == L_IDLE && i < n) {
					slot = b.qVol[i];
					float4 A = b.rec[slot].pA, B = b.rec[slot].pB, C = b.rec[slot].pC;
					bg = b.rec[slot].pD.x;
					float tNear = b.rec[slot].hA.w;
					tFar = b.rec[slot].hB.w - tNear;  // volume_enter, Li :198-201
					int inst = __float_as_int(b.rec[slot].hC.z);
					ray.o = V3(A.x, A.y, A.z);
					ray.d = V3(A.w, B.x, B.y);
					ray.o = ray.at(tNear);
					const DInstance& in = S.inst[inst];
					const DMaterial& m = S.mat[in.material];
					vol = &S.vol[m.volume];
					rng.init(P.seed, __float_as_uint(C.y), __float_as_uint(C.z), __float_as_uint(C.w));
					wr.start(rng);
					trk.init(*vol, m, transform_ray(ray, in.Mi), 0.0f, tFar, wr, st);
					budget = P.budget;
					state = L_MOVING;
				}
				if (__ballot_sync(0xffffffffu, state == L_IDLE && i >= n)) exhausted = true;
			}
			if (__ballot_sync(0xffffffffu, state != L_IDLE) == 0) break;  // nothing walking, nothing left to fetch
		}
		// ---- move
#pragma unroll 1
		for (int k = 0; k < P.moves; k++) {
			if (state == L_MOVING) {
				if (budget-- <= 0) state = L_FIN_BUDGET;
				else if (trk.wants_candidate(wr)) state = L_CAND;
				else if (trk.move(st) == TRACK_END) state = L_FIN_END;
			}
			if (NE_TRACK_EARLY_BREAK && !__any_sync(0xffffffffu, state == L_MOVING)) break;
		}
		// ---- candidate
		if (state == L_CAND) {
			float density = trk.candidate_density(*vol);
			state = delta_candidate(trk, density, wr, st) == TRACK_CANDIDATE ? L_FIN_HIT : L_MOVING;
		}
	}
	flush_stats_wf(st, P.counters);
}

// Real collisions: phase function, next-event setup, continuation (Li :215-236).
// FUSE (scenes without meshes, where intersectScene is a handful of analytic tests): the scattered ray is traced and
// classified right here, so a path that goes on through a grid medium enters the NEXT iteration's volume queue
// directly, one that reaches a surface enters this iteration's surface queue, and one that ends frees its slot: no
// trip through the extend queue, i.e. one 128-byte record read and one hit write fewer per scatter event.
template <bool FUSE>
__global__ void __launch_bounds__(256) k_wf_scatter(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t n = b.c->scat;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qScat[i];
		PathRec r = load_path(b, slot);
		Hit h = load_hit(b, slot);
		PhiloxRng rng;
		rng.init(P.seed, r.pixel, r.sample, r.dim);
		QueueSink sink;
		sink.b = &b;
		sink.accum = P.accum;
		sink.pixel = r.pixel;
		sink.sample = r.sample;
		int next;
		if (S.mat[S.inst[h.inst].material].volume < 0) {
			next = shade_volume_homog(S, r.ps, h, rng, sink, st);
		} else {
			Ray rayO = transform_ray(r.ps.ray, S.inst[h.inst].Mi);
			next = volume_scatter(S, r.ps, h, rayO, r.tHit, rng, sink, st);
		}
		if (next == PATH_NEXT_BOUNCE) r.ps.bounce++;
		if (next == PATH_DONE || r.ps.bounce >= P.bounces) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			r.dim = rng.dim;
			if (FUSE) {
				Hit h2;
				const uint32_t prims0 = st.prim_tests;
				bool did = intersect_scene_nomesh(S, r.ps.ray, h2, float(NE_EPSILON12), INFINITY, st);
				int kind = classify_hit(S, did, h2, r.ps, sink);
				const bool grid = kind == HIT_VOLUME && S.mat[S.inst[h2.inst].material].volume >= 0;
				if (kind == HIT_VOLUME && !grid) {
					// a HomogeneousMedia hit belongs in the scatter queue, which this kernel is draining: leave it to k_wf_extend
					st.prim_tests = prims0;  // k_wf_extend will count the query
					store_path(b, slot, r);
					b.qNext[warp_push(&b.c->next)] = slot;
					continue;
				}
				st.extend_rays++;
				if (kind == HIT_TERMINATE) {
					b.qFree[warp_push(&b.c->freeN)] = slot;
					continue;
				}
				store_path(b, slot, r);
				store_hit(b, slot, h2);
				if (grid) b.qVolNext[warp_push(&b.c->volNext)] = slot;
				else b.qSurf[warp_push(&b.c->surf)] = slot;
				continue;
			}
			store_path(b, slot, r);
			b.qNext[warp_push(&b.c->next)] = slot;
		}
	}
	flush_stats_wf(st, P.counters);
}

// Surface hits: GGX shading, next-event setup, continuation (Li :262-283).
__global__ void __launch_bounds__(256) k_wf_surface(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t n = b.c->surf;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qSurf[i];
		PathRec r = load_path(b, slot);
		Hit h = load_hit(b, slot);
		PhiloxRng rng;
		rng.init(P.seed, r.pixel, r.sample, r.dim);
		QueueSink sink;
		sink.b = &b;
		sink.accum = P.accum;
		sink.pixel = r.pixel;
		sink.sample = r.sample;
		int next = shade_surface<PhiloxRng>(S, r.ps, h, rng, sink, st);
		if (next == PATH_NEXT_BOUNCE) r.ps.bounce++;
		if (next == PATH_DONE || r.ps.bounce >= P.bounces) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			r.dim = rng.dim;
			store_path(b, slot, r);
			b.qNext[warp_push(&b.c->next)] = slot;
		}
	}
	flush_stats_wf(st, P.counters);
}

// visibilityTr requests: splat the weight when nothing or an emitter is hit first.
__global__ void __launch_bounds__(256) k_wf_shadow(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t n = b.c->shadow;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 A = b.sA[i], B = b.sB[i];
		float2 C = b.sC[i];
		PhiloxRng dummy;
		float vis = visibility_tr<PhiloxRng, false, true>(S, V3(A.x, A.y, A.z), V3(A.w, B.x, B.y), dummy, st);
		if (vis != 0) splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * vis);
	}
	flush_stats_wf(st, P.counters);
}

// intersectTr :13-31 for the requests pushed in this iteration: march THROUGH non-medium surfaces until a medium
// (the request gets its instance, entry point and segment length) or nothing (the request is dropped: Li = 0).
__global__ void __launch_bounds__(256) k_wf_trfind(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t first = b.c->trNew0, n = b.c->tr;
	Stats st;
	st.clear();
	for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 A = b.tA[i], B = b.tB[i];
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y);
		int inst = -2;
		float tRemain = 0;
		for (int seg = 0; seg < NE_MAX_TR_SEGMENTS; seg++) {
			Hit hh;
			st.shadow_rays++;
			if (!intersect_scene(S, ray, hh, float(NE_EPSILON3), INFINITY, st)) break;
			int mi = S.inst[hh.inst].material;
			if (mi >= 0 && S.mat[mi].has_medium && S.mat[mi].volume >= 0) {
				inst = hh.inst;
				ray.o = ray.at(hh.tNear);
				tRemain = hh.tFar - hh.tNear;
				break;
			}
			if (mi >= 0 && S.mat[mi].has_medium) {  // HomogeneousMedia: closed-form transmittance, nothing to walk
				float4 C = b.tC[i];
				splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * homog_tr(S.mat[mi], hh.tFar - hh.tNear));
				break;  // inst stays -2: the request is done
			}
			ray.o = hh.p;
		}
		b.tA[i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
		b.tD[i] = make_float4(1.0f, tRemain, __int_as_float(inst), __uint_as_float(0u));
	}
	flush_stats_wf(st, P.counters);
}

// ---------------------------------------------------------------------------------------------------------------
// Scenes WITH triangle meshes: the three ray-casting stages (extend, shadow, trfind) as PERSISTENT warps over the
// resumable SceneTrace. Traversal lengths of incoherent rays are heavy-tailed; in the grid-stride kernels above a warp
// lasts as long as its longest ray (ncu on C4: 7 of 32 lanes active). Here a warp alternates two warp-wide phases:
//   refill   (when P.walkRefill lanes are not walking) lanes whose BVH walk ended complete their mesh instance and go
//            on with the instance fold; lanes whose fold ended hand the answer to the job (classify / splat / locate
//            the medium) and fetch the next ray (one atomicAdd per warp); new rays fold up to their first mesh
//   walk     every walking lane spends at most P.walkBudget inner-node visits of the while-while traversal
// Each ray still sees intersect_scene's exact sequence of operations. Scenes without meshes keep the kernels above
// (their fold is a handful of analytic tests: nothing to balance).
// ---------------------------------------------------------------------------------------------------------------
enum { T_IDLE = 0, T_FOLD = 1, T_WALK = 2, T_WALKED = 3, T_FOLDED = 4 };

// k_wf_extend's job: Scene::intersectScene for the extend queue + classify (Li :187-193, :244-260).
struct ExtendJob {
	uint32_t slot, pixel;
	V3 T;
	int bounce;
	__device__ __forceinline__ static uint32_t* head(const WfBuf& b) { return &b.c->extHead; }
	__device__ __forceinline__ static uint32_t first(const WfBuf&) { return 0; }
	__device__ __forceinline__ static uint32_t count(const WfBuf& b) { return b.c->extend; }
	__device__ __forceinline__ void fetch(const WfBuf& b, uint32_t i, SceneTrace& tr, Stats& st) {
		slot = b.qExtend[i];
		float4 A = b.rec[slot].pA, B = b.rec[slot].pB, C = b.rec[slot].pC;
		bounce = int(b.rec[slot].pD.x & 0xff);
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y);
		T = V3(B.z, B.w, C.x);
		pixel = __float_as_uint(C.y);
		st.extend_rays++;
		tr.begin(ray, float(NE_EPSILON12), INFINITY);
	}
	__device__ __forceinline__ bool consume(const DScene& S, const WfBuf& b, const WfParams& P, SceneTrace& tr, Stats&) {
		PathState ps;
		ps.ray = tr.rayW;
		ps.T = T;
		ps.bounce = bounce;
		QueueSink sink;
		sink.accum = P.accum;
		sink.pixel = pixel;
		int kind = classify_hit(S, tr.did, tr.hit, ps, sink);
		if (kind == HIT_TERMINATE) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			store_hit(b, slot, tr.hit);
			if (kind == HIT_VOLUME) {
				if (S.mat[S.inst[tr.hit.inst].material].volume >= 0) b.qVol[warp_push(&b.c->vol)] = slot;
				else b.qScat[warp_push(&b.c->scat)] = slot;
			} else b.qSurf[warp_push(&b.c->surf)] = slot;
		}
		return false;
	}
};

// k_wf_shadow's job: visibilityTr :34-72 for the shadow requests.
struct ShadowJob {
	uint32_t req;
	__device__ __forceinline__ static uint32_t* head(const WfBuf& b) { return &b.c->shHead; }
	__device__ __forceinline__ static uint32_t first(const WfBuf&) { return 0; }
	__device__ __forceinline__ static uint32_t count(const WfBuf& b) { return b.c->shadow; }
	__device__ __forceinline__ void fetch(const WfBuf& b, uint32_t i, SceneTrace& tr, Stats& st) {
		req = i;
		float4 A = b.sA[i], B = b.sB[i];
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y) - ray.o;
		st.shadow_rays++;
		tr.begin(ray, float(NE_EPSILON3), INFINITY);
	}
	__device__ __forceinline__ bool consume(const DScene& S, const WfBuf& b, const WfParams& P, SceneTrace& tr, Stats&) {
		bool vis = true;
		if (tr.did) {
			int mi = S.inst[tr.hit.inst].material;
			vis = mi >= 0 && S.mat[mi].has_light;
		}
		if (vis) {
			float4 B = b.sB[req];
			float2 C = b.sC[req];
			splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x));
		}
		return false;
	}
};

// k_wf_trfind's job: intersectTr :13-31 for the requests pushed in this iteration.
struct TrFindJob {
	uint32_t req;
	int seg;
	__device__ __forceinline__ static uint32_t* head(const WfBuf& b) { return &b.c->trfHead; }
	__device__ __forceinline__ static uint32_t first(const WfBuf& b) { return b.c->trNew0; }
	__device__ __forceinline__ static uint32_t count(const WfBuf& b) { return b.c->tr - b.c->trNew0; }
	__device__ __forceinline__ void fetch(const WfBuf& b, uint32_t i, SceneTrace& tr, Stats& st) {
		req = i;
		seg = 0;
		float4 A = b.tA[i], B = b.tB[i];
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y);
		st.shadow_rays++;
		tr.begin(ray, float(NE_EPSILON3), INFINITY);
	}
	__device__ __forceinline__ bool consume(const DScene& S, const WfBuf& b, const WfParams& P, SceneTrace& tr, Stats& st) {
		Ray ray = tr.rayW;
		int inst = -2;
		float tRemain = 0;
		if (tr.did) {
			const Hit& hh = tr.hit;
			int mi = S.inst[hh.inst].material;
			if (mi >= 0 && S.mat[mi].has_medium && S.mat[mi].volume >= 0) {
				inst = hh.inst;
				ray.o = ray.at(hh.tNear);
				tRemain = hh.tFar - hh.tNear;
			} else if (mi >= 0 && S.mat[mi].has_medium) {  // HomogeneousMedia: closed-form transmittance, nothing to walk
				float4 B = b.tB[req], C = b.tC[req];
				splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * homog_tr(S.mat[mi], hh.tFar - hh.tNear));
			} else {
				ray.o = hh.p;
				if (++seg < NE_MAX_TR_SEGMENTS) {  // through the surface: next segment of the same request
					st.shadow_rays++;
					tr.begin(ray, float(NE_EPSILON3), INFINITY);
					return true;
				}
			}
		}
		b.tA[req] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
		b.tD[req] = make_float4(1.0f, tRemain, __int_as_float(inst), __uint_as_float(0u));
		return false;
	}
};

#ifndef NE_TRACE_BLOCKS
#define NE_TRACE_BLOCKS 4  // resident 256-thread blocks per SM the trace kernels are compiled for (64 registers: the walk is bound by
                           // the latency of dependent node loads, so warps in flight count; measured 2: 43.5, 3: 33.9, 4: 31.0, 5: 33.3 ms on C4)
#endif
template <class JOB>
__global__ void __launch_bounds__(256, NE_TRACE_BLOCKS) k_wf_trace(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	const uint32_t first = JOB::first(b), n = JOB::count(b);
	Stats st;
	st.clear();
	JOB job;
	SceneTrace tr;
	int stack[NE_BVH_STACK];
	int state = T_IDLE;
	bool exhausted = false;
	while (true) {
		const unsigned walking = __ballot_sync(0xffffffffu, state == T_WALK);
		const bool pending = !exhausted || __any_sync(0xffffffffu, state == T_WALKED);  // something a refill phase could do
		if (walking == 0 || (pending && 32 - __popc(walking) >= P.walkRefill)) {
			// ---- refill: runs until every lane either walks a BVH or has nothing left to do
			while (true) {
				if (state == T_WALKED) {
					tr.mesh_done(S);
					state = T_FOLD;
				}
				if (state == T_FOLDED) state = job.consume(S, b, P, tr, st) ? T_FOLD : T_IDLE;
				if (!exhausted) {
					WarpReserve rf;
					rf.issue(JOB::head(b), state == T_IDLE);
					const uint32_t i = rf.get();
					if (state == T_IDLE && i < n) {
						job.fetch(b, first + i, tr, st);
						state = T_FOLD;
					}
					if (__ballot_sync(0xffffffffu, state == T_IDLE)) exhausted = true;  // a lane came back empty-handed
				}
				if (state == T_FOLD) state = tr.fold(S, stack, st) ? T_WALK : T_FOLDED;
				if (!__any_sync(0xffffffffu, state == T_FOLDED)) break;  // T_FOLD / T_WALKED cannot be pending here
			}
			if (__ballot_sync(0xffffffffu, state != T_IDLE) == 0) break;
		}
		// ---- walk
		if (state == T_WALK) {
			const DMesh& m = S.mesh[S.inst[tr.i].mesh];
			if (tr.walk(m, stack, P.walkBudget, st)) state = T_WALKED;
		}
	}
	flush_stats_wf(st, P.counters);
}

// Transmittance requests whose medium is known: ratio tracking through it (at most P.budget events per pass), splat
// weight * Tr. Persistent warps and phases like k_wf_track; the weight and pixel are re-read from the request when the
// walk ends.
template <int BRICKMAJ>
__global__ void __launch_bounds__(NE_TRACK_THREADS, NE_TRACK_BLOCKS) k_wf_tr(WfBuf b, WfParams P) {
	NE_STAGE_SCENE();
	typedef typename WalkRngOf<(BRICKMAJ != 0), PhiloxRng>::type WalkRng;
	const uint32_t n = b.c->tr;
	Stats st;
	st.clear();
	int state = L_IDLE;
	bool exhausted = false;
	uint32_t req = 0;
	Ray ray;  // WCS, origin at the start of the remaining segment
	float Tr = 1, tRemain = 0;
	int inst = -1;
	PhiloxRng rng;
	WalkRng wr;
	Tracker<BRICKMAJ> trk;
	const DVolume* vol = nullptr;
	int budget = 0;
	while (true) {
		unsigned walking = __ballot_sync(0xffffffffu, state == L_MOVING || state == L_CAND);
		if (walking == 0 || (!exhausted && 32 - __popc(walking) >= P.refill)) {
			// ---- finish + refill
			const bool fin = state >= L_FIN_HIT;
			WarpReserve rNext, rFetch;
			rNext.issue(&b.c->trNext, state == L_FIN_BUDGET);
			rFetch.issue(&b.c->trHead, !exhausted && (state == L_IDLE || fin));
			const uint32_t iNext = rNext.get(), iFetch = rFetch.get();  // warp-wide shuffles: before the lanes part ways
			if (fin) {
				float4 B = b.tB[req], C = b.tC[req];
				if (state == L_FIN_BUDGET) {
					uint32_t j = iNext;
					V3 o = ray.at(trk.t);
					b.uA[j] = make_float4(o.x, o.y, o.z, ray.d.x);
					b.uB[j] = B;
					b.uC[j] = C;
					b.uD[j] = make_float4(Tr, tRemain - trk.t, __int_as_float(inst), __uint_as_float(rng.dim));
				} else if (Tr != 0) {
					splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * V3(Tr));
				}
				state = L_IDLE;
			}
			if (!exhausted) {
				const uint32_t i = iFetch;
				if (state == L_IDLE && i < n) {
					float4 D = b.tD[i];
					inst = __float_as_int(D.z);
					if (inst >= 0) {  // else no medium along the ray: the request is dropped
						float4 A = b.tA[i], B = b.tB[i], C = b.tC[i];
						req = i;
						ray.o = V3(A.x, A.y, A.z);
						ray.d = V3(A.w, B.x, B.y);
						Tr = D.x;
						tRemain = D.y;
						rng.init(P.seed, __float_as_uint(C.y), __float_as_uint(C.z), __float_as_uint(D.w), __float_as_uint(C.w));
						const DInstance& in = S.inst[inst];
						const DMaterial& m = S.mat[in.material];
						vol = &S.vol[m.volume];
						wr.start(rng);
						trk.init(*vol, m, transform_ray(ray, in.Mi), 0.0f, tRemain, wr, st);  // GridMedia::Tr :49
						budget = P.budget;
						state = L_MOVING;
					}
				}
				if (__ballot_sync(0xffffffffu, state == L_IDLE && i >= n)) exhausted = true;
			}
			if (__ballot_sync(0xffffffffu, state != L_IDLE) == 0) {
				if (exhausted) break;
				continue;  // every request of this batch was a dropped one: fetch again
			}
		}
		// ---- move
#pragma unroll 1
		for (int k = 0; k < P.moves; k++) {
			if (state == L_MOVING) {
				if (budget-- <= 0) state = L_FIN_BUDGET;
				else if (trk.wants_candidate(wr)) state = L_CAND;
				else if (trk.move(st) == TRACK_END) state = L_FIN_END;
			}
			if (NE_TRACK_EARLY_BREAK && !__any_sync(0xffffffffu, state == L_MOVING)) break;
		}
		// ---- candidate
		if (state == L_CAND) {
			float density = trk.candidate_density(*vol);
			state = ratio_candidate(trk, density, Tr, wr, st) == TRACK_END ? L_FIN_END : L_MOVING;
		}
	}
	flush_stats_wf(st, P.counters);
}

}  // namespace

struct ne_wavefront_state {
	WfBuf b{};
	uint32_t nSlots = 0;
	std::vector<void*> allocs;
	uint32_t* hostDone = nullptr;  // pinned, mapped
	uint32_t* devDone = nullptr;
	int gridBlocks = 0, smCount = 0;
	std::vector<cudaEvent_t> events;
};

namespace ne {

void wavefront_free(ne_b200_ctx* ctx) {
	ne_wavefront_state* w = ctx->wf;
	if (!w) return;
	for (void* p : w->allocs) cudaFree(p);
	if (w->hostDone) cudaFreeHost(w->hostDone);
	for (cudaEvent_t e : w->events) cudaEventDestroy(e);
	delete w;
	ctx->wf = nullptr;
}

template <class T>
static int wf_alloc(ne_wavefront_state* w, T** p, size_t n) {
	NE_CUDA_OK(cudaMalloc(p, n * sizeof(T)));
	w->allocs.push_back(*p);
	return NE_B200_OK;
}

static int wavefront_ensure(ne_b200_ctx* ctx, uint32_t nSlots) {
	if (ctx->wf && ctx->wf->nSlots >= nSlots) return NE_B200_OK;
	wavefront_free(ctx);
	ne_wavefront_state* w = new ne_wavefront_state();
	ctx->wf = w;
	w->nSlots = nSlots;
	int rc;
	WfBuf& b = w->b;
#define A(field) if ((rc = wf_alloc(w, &b.field, nSlots))) return rc;
	A(rec) A(sA) A(sB) A(sC) A(tA) A(tB) A(tC) A(tD) A(uA) A(uB) A(uC) A(uD)
	A(qExtend) A(qNext) A(qVol) A(qVolNext) A(qScat) A(qSurf) A(qFree)
#undef A
	if ((rc = wf_alloc(w, &b.c, 1))) return rc;
	NE_CUDA_OK(cudaHostAlloc(&w->hostDone, sizeof(uint32_t), cudaHostAllocMapped));
	NE_CUDA_OK(cudaHostGetDevicePointer(&w->devDone, w->hostDone, 0));
	cudaDeviceProp prop;
	NE_CUDA_OK(cudaGetDeviceProperties(&prop, ctx->device));
	w->smCount = prop.multiProcessorCount;
	w->gridBlocks = prop.multiProcessorCount * 8;  // 148 SMs x 8 resident 256-thread blocks
	return NE_B200_OK;
}

static uint32_t env_u32(const char* name, uint32_t dflt) {
	const char* e = getenv(name);
	return e ? (uint32_t)strtoul(e, nullptr, 10) : dflt;
}

int wavefront_render(ne_b200_ctx* ctx, int sppBegin, int sppEnd, int bounces, uint64_t seed, uint32_t flags) {
	const unsigned long long work = (unsigned long long)ctx->W * ctx->H * (unsigned long long)(sppEnd - sppBegin);
	if (work == 0 || bounces == 0) return NE_B200_OK;
	uint32_t pool = std::max(1024u, env_u32("NE_B200_POOL", 1u << 26));
	uint32_t nSlots = uint32_t(std::min<unsigned long long>(work, pool));
	int rc = wavefront_ensure(ctx, nSlots);
	if (rc) return rc;
	ne_wavefront_state* w = ctx->wf;
	nSlots = w->nSlots;
	WfParams P;
	P.s = ctx->scene;
	P.cam = ctx->cam;
	P.accum = ctx->accum;
	P.W = ctx->W;
	P.H = ctx->H;
	P.sppBegin = sppBegin;
	P.bounces = bounces;
	P.budget = int(std::max(1u, env_u32("NE_B200_TRACK_BUDGET", 64)));
	P.moves = int(std::max(1u, env_u32("NE_B200_TRACK_MOVES", 4)));  // brick crossings per lane between two candidate phases
	P.refill = int(std::min(32u, std::max(1u, env_u32("NE_B200_TRACK_REFILL", 20))));  // refill a warp once this many lanes are idle
	P.walkBudget = int(std::max(1u, env_u32("NE_B200_WALK_BUDGET", 24)));  // inner-node visits per walking lane between two votes
	P.walkRefill = int(std::min(32u, std::max(1u, env_u32("NE_B200_WALK_REFILL", 12))));  // refill a trace warp once this many lanes are not walking
	P.seed = seed;
	P.counters = ctx->dCounters;
	// persistent trace kernels when there are BVHs to walk (NE_B200_TRACE=0/1 overrides)
	const bool trace = env_u32("NE_B200_TRACE", ctx->nMeshes > 0 ? 1 : 0) != 0;
	const int GR = w->smCount * NE_TRACE_BLOCKS;
	// without meshes, k_wf_scatter traces its own continuation ray (NE_B200_FUSE=0/1 overrides)
	const bool fuse = ctx->nMeshes == 0 && env_u32("NE_B200_FUSE", 1) != 0;
	const bool brick = !(flags & NE_B200_RENDER_GLOBAL_MAJORANT);
	const int genBlocks = int(env_u32("NE_B200_GEN_BLOCKS", trace ? 4 : 3));  // resident blocks k_wf_generate is compiled for (C2: 2: 33.9, 3: 33.2, 4: 33.6 ms)
	// empty-space skipping in the tracking walks pays where a good part of a brick table is far from any density
	// (ctx->skipWorthwhile, decided at upload; NE_B200_SKIP=0/1 overrides)
	const bool skip = env_u32("NE_B200_SKIP", ctx->skipWorthwhile ? 1 : 0) != 0;
	cudaStream_t st = ctx->stream;
	const int G = w->gridBlocks, B = 256;
	const int GT = w->smCount * NE_TRACK_BLOCKS;  // persistent tracking kernels: exactly the resident blocks

	// event pool for per-stage device times, resolved after every host poll
	size_t evUsed = 0;
	auto ev = [&]() -> cudaEvent_t {
		if (evUsed == w->events.size()) {
			cudaEvent_t e;
			cudaEventCreate(&e);
			w->events.push_back(e);
		}
		cudaEvent_t e = w->events[evUsed++];
		cudaEventRecord(e, st);
		return e;
	};
	struct Span { cudaEvent_t a, b; int kind; };
	std::vector<Span> spans;
	const bool timeStages = getenv("NE_B200_NO_STAGE_TIMES") == nullptr;

	k_wf_init<<<(nSlots + 255) / 256, 256, 0, st>>>(w->b, nSlots, work);
	ctx->kernelLaunches++;
	*w->hostDone = 0;
	bool done = false;
	unsigned long long iter = 0;
	while (!done) {
		// a few iterations per host poll; finished iterations cost only empty launches
		for (int k = 0; k < 4; k++) {
			WfBuf b = w->b;
			if (iter & 1) {
				std::swap(b.qExtend, b.qNext);
				std::swap(b.qVol, b.qVolNext);
				std::swap(b.tA, b.uA);
				std::swap(b.tB, b.uB);
				std::swap(b.tC, b.uC);
				std::swap(b.tD, b.uD);
			}
			k_wf_plan<<<1, 1, 0, st>>>(b, w->devDone);
			// camera rays are coherent: the grid-stride kernel is as fast as a trace job (measured)
			if (genBlocks >= 4) k_wf_generate<4><<<G, B, 0, st>>>(b, P);
			else if (genBlocks == 3) k_wf_generate<3><<<G, B, 0, st>>>(b, P);
			else k_wf_generate<2><<<G, B, 0, st>>>(b, P);
			k_wf_commit<<<1, 1, 0, st>>>(b, ctx->dCounters);
			cudaEvent_t e0 = timeStages ? ev() : nullptr;
			if (trace) k_wf_trace<ExtendJob><<<GR, 256, 0, st>>>(b, P);
			else k_wf_extend<<<G, B, 0, st>>>(b, P);
			cudaEvent_t e1 = timeStages ? ev() : nullptr;
			if (!brick) k_wf_track<TRACK_GLOBAL><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
			else if (skip) k_wf_track<TRACK_SKIP><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
			else k_wf_track<TRACK_BRICK><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
			cudaEvent_t e2 = timeStages ? ev() : nullptr;
			if (fuse) k_wf_scatter<true><<<G, B, 0, st>>>(b, P);
			else k_wf_scatter<false><<<G, B, 0, st>>>(b, P);
			k_wf_surface<<<G, B, 0, st>>>(b, P);
			cudaEvent_t e3 = timeStages ? ev() : nullptr;
			if (trace) {
				k_wf_trace<ShadowJob><<<GR, 256, 0, st>>>(b, P);
				k_wf_trace<TrFindJob><<<GR, 256, 0, st>>>(b, P);
			} else {
				k_wf_shadow<<<G, B, 0, st>>>(b, P);
				k_wf_trfind<<<G, B, 0, st>>>(b, P);
			}
			cudaEvent_t e4 = timeStages ? ev() : nullptr;
			if (!brick) k_wf_tr<TRACK_GLOBAL><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
			else if (skip) k_wf_tr<TRACK_SKIP><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
			else k_wf_tr<TRACK_BRICK><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
			cudaEvent_t e5 = timeStages ? ev() : nullptr;
			if (timeStages) {
				spans.push_back({e0, e1, 0});  // extend
				spans.push_back({e1, e2, 1});  // delta tracking
				spans.push_back({e2, e3, 2});  // scatter + surface shading
				spans.push_back({e3, e4, 0});  // shadow rays
				spans.push_back({e4, e5, 1});  // ratio tracking
			}
			ctx->kernelLaunches += 10;
			iter++;
		}
		NE_CUDA_OK(cudaStreamSynchronize(st));
		NE_CUDA_OK(cudaGetLastError());
		done = *w->hostDone != 0;
		if (timeStages) {
			for (const Span& s : spans) {
				float ms = 0;
				cudaEventElapsedTime(&ms, s.a, s.b);
				(s.kind == 0 ? ctx->msExtend : s.kind == 1 ? ctx->msVolume : ctx->msShade) += ms;
			}
			spans.clear();
			evUsed = 0;
		}
	}
	ctx->wavefrontIterations += iter;
	return NE_B200_OK;
}

}  // namespace ne
