// ne_wavefront.cu — the production renderer: OfflineEngine::renderTile's pixel x sample loops
// (core/OfflineEngine.cpp:61-71) over the whole frame as a WAVEFRONT of path records.
//
//   pool      N path slots (one 128-byte record each: 64 B path state + 48 B hit) that are refilled with new camera
//             samples as paths terminate, so the wavefront stays full until the work runs out. 64 Mi slots by default
//             (HBM is plentiful, and a wide wavefront means few, long launches); never initialised: slots come from a
//             bump counter until the first ones return through the free stack
//   queues    arrays of slot indices: {generate, extend} -> {volume, scatter (homogeneous media), surface};
//             volume -> {scatter, volume (walk not finished), next extend}; free slots. Pushes are warp-aggregated
//             (one atomicAdd per warp per queue; the persistent tracking kernels reserve theirs in chunks, WarpChunk)
//   kernels   init / plan / commit / finish (1 thread: bookkeeping; plan also sets the render graph's loop condition)
//             · generate (camera ray + the path's first Scene::intersectScene, for the pixels of the culling rectangle
//               only: paths that end at once never enter the pool)
//             · extend (Scene::intersectScene fold, BVH, for continuing paths)
//             · track (delta tracking: persistent warps in finish+refill / move / candidate phases)
//             · scatter (phase function + next-event setup) · surface (GGX shading + next-event setup)
//             · shadow (visibilityTr requests) · trfind (intersectTr: walk through surfaces to the first medium)
//             · tr (ratio tracking through that medium, persistent warps like track)
//             · trace<Extend|Shadow|TrFind job> (scenes with meshes: the three ray-casting stages as persistent warps over
//               a resumable intersectScene, so a warp is not held by its longest BVH walk)
//             variants chosen per scene by the host: scatter<FUSE> traces its own continuation ray in mesh-free scenes;
//             scatter / surface<LS> are specialised for the scene's light set; track/tr<TRACK_*_SM> keep the majorant tables
//             in shared memory (bulk async copy) when they fit, <TRACK_SKIP*> cross cubes of empty bricks in one move
//   driver    a whole render is ONE CUDA graph: first plan -> WHILE(not done){ iteration; plan } -> finish, the loop condition set
//             on the device (graph_build); ne_b200_render is asynchronous. Large batches run as two such graphs on two streams
//             (wavefront_render). NE_B200_HOST_LOOP=1: the same kernels launched one by one with per-stage CUDA events.
//   output    fp32 atomicAdd splats into the context's linear accumulation buffer
//
// The two tracking kernels stop a walk after `budget` events (brick crossings + density look-ups) and queue the remainder
// for the next pass: exponential free flights are memoryless, so the estimate is unchanged, and a launch is not held hostage
// by its longest walks. Every kernel runs over a device-side count with a fixed
// grid of (SM count x resident blocks): no host round trip sizes a launch.
// Every block stages the scene's instance / material / volume tables in shared memory first (stage_scene).
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ne_ctx.h"
#include "ne_integrator.cuh"

using namespace ne;

namespace {

// Per-render arguments. The renderer's launches live in a CUDA graph whose kernel parameters are fixed when it is built,
// so what changes from one ne_b200_render to the next is written here by k_wf_init and read from global memory.
struct WfDyn {
	DCamera cam;
	float* accum;
	unsigned long long seed;
	int sppBegin, bounces;
	// camera rays are generated for the pixels of this rectangle only (cull_rect: everything hittable projects inside it)
	int rx0, ry0, rw, rh;
};

enum { STAGE_TRACE = 0, STAGE_VOLUME = 1, STAGE_SHADE = 2, STAGE_OTHER = 3, STAGE_KINDS = 4 };

struct WfCounts {
	uint32_t extend, next, vol, volNext, volHead, scat, surf, freeN, shadow, tr, trNext, trHead, trNew0, gen, genTaken, done, extHead, shHead, trfHead;
	uint32_t par;       // which buffer of each ping-pong pair is "current" (flipped by k_wf_plan)
	uint32_t bump;      // slots >= bump have never been handed out: the pool needs no initialisation pass
	uint32_t bumpBase, freeTake;  // this iteration's refill: freeTake slots off the free stack, the rest from bumpBase on
	uint32_t overflow;  // a request array was full (cannot happen by construction; checked so it could never go unnoticed)
	unsigned long long workNext, workTotal;
	unsigned long long lastNs, stageNs[STAGE_KINDS], iterations, launches;
	WfDyn dyn;
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
	//   path  pA=(o.xyz,d.x) pB=(d.yz,T.xy) pC=(T.z,pixel,sample,dim) pD=(bounce|guard<<16, nee, collision t, -)
	//   hit   hA=(p.xyz,tNear) hB=(n.xyz,tFar) hC=(u,v,inst,prim)
	PathSlot* rec;
	// shadow request: A=(o.xyz,C.x) B=(C.yz,w.xy) C=(w.z,pixel)
	float4 *sA, *sB;
	float2* sC;
	// transmittance request, [par] = this pass, [par ^ 1] = next pass: A=(o.xyz,d.x) B=(d.yz,w.xy) C=(w.z,pixel,sample,stream)
	// D=(Tr so far, remaining tFar, medium instance or -1 = not found yet, dim)
	float4 *tA[2], *tB[2], *tC[2], *tD[2];
	// index queues; [par] = this iteration's, [par ^ 1] = the next one's
	uint32_t *qExt[2], *qVol[2];
	uint32_t *qScat, *qSurf, *qFree;
	WfCounts* c;
	// capacities: every live slot sits in exactly one stage queue and is shaded at most once per iteration, so one
	// iteration pushes at most nSlots shadow and nSlots new transmittance requests; walks cut by the event budget are
	// carried over only while there is room (else they simply keep walking), so trCap = nSlots + carryCap is never exceeded
	uint32_t nSlots, shadowCap, trCap, carryCap;
};

struct WfParams {
	DScene s;
	int W, H, budget, refill, moves, walkBudget, walkRefill, cutAlways;
	int genToVol;   // every camera path that survives goes to the volume queue (ctx->onlyGridMedia): its queue entry is its reservation
	int prevStage;  // stage kind of the kernel launched before this one, or -1: no stage accounting (see stage_stamp)
	DCounters* counters;
};

// Block barrier after code in which the lanes of a warp may have parted ways (a loop with per-thread trip counts, work done
// by thread 0 only): the warp reconverges first, so every lane arrives at the barrier together (compute-sanitizer synccheck).
__device__ __forceinline__ void block_sync() {
	__syncwarp();
	__syncthreads();
}

#define NE_QCHUNK 64u             // entries a tracking warp reserves at a time in its hot queues (WarpChunk)
#define NE_Q_INVALID 0xffffffffu  // a queue entry that holds no slot (the unused tail of a warp's last chunk): consumers skip it

__device__ __forceinline__ uint32_t warp_push(uint32_t* counter) {
	unsigned m = __activemask();
	unsigned lane = threadIdx.x & 31;
	int leader = __ffs(m) - 1;
	uint32_t base = 0;
	if (int(lane) == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
	base = __shfl_sync(m, base, leader);
	return base + __popc(m & ((1u << lane) - 1));
}

// Work counters: per thread -> warp (one reduction instruction per field) -> block (shared-memory atomics) -> ONE global
// atomic per field and block. Every warp used to add its totals to the ten global counters itself: thousands of
// same-address atomics at the end of every launch, serialised in L2 - 10 to 20 microseconds per kernel, as much as the
// work itself in the late, nearly empty iterations of a frame (found in the 8-spp-per-GPU frames of the strong-scaling run).
// Must be reached by EVERY thread of the block (it synchronises).
#define NE_STAT_FIELDS 10
__device__ __forceinline__ void flush_stats_wf(const Stats& st, DCounters* c) {
	__shared__ unsigned long long blockTotals_[NE_STAT_FIELDS];
	if (threadIdx.x < NE_STAT_FIELDS) blockTotals_[threadIdx.x] = 0;
	block_sync();
	const unsigned lane = threadIdx.x & 31;
#define NE_FLUSH(k, field)                                                          \
	{                                                                               \
		unsigned w = __reduce_add_sync(0xffffffffu, (unsigned)(st.field));          \
		if (lane == 0 && w) atomicAdd(&blockTotals_[k], (unsigned long long)w);     \
	}
	NE_FLUSH(0, extend_rays) NE_FLUSH(1, shadow_rays) NE_FLUSH(2, delta_steps) NE_FLUSH(3, ratio_steps) NE_FLUSH(4, brick_visits)
	NE_FLUSH(5, bvh_nodes) NE_FLUSH(6, tri_tests) NE_FLUSH(7, prim_tests) NE_FLUSH(8, scatter_events) NE_FLUSH(9, surface_events)
#undef NE_FLUSH
	block_sync();
	// the ten counters are consecutive 64-bit fields of DCounters, in this order, after `paths`
	if (threadIdx.x < NE_STAT_FIELDS && blockTotals_[threadIdx.x]) atomicAdd(&c->extend_rays + threadIdx.x, blockTotals_[threadIdx.x]);
}
static_assert(offsetof(DCounters, surface_events) - offsetof(DCounters, extend_rays) == 9 * sizeof(unsigned long long), "DCounters field order");

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
	r.ps.bounce = int(D.x & 0xffffu);  // 16 bits each: ne_b200_render rejects more than 65535 bounces
	r.ps.guard = int(D.x >> 16);
	r.ps.nee = D.y;
	r.tHit = __uint_as_float(D.z);
	return r;
}
__device__ __forceinline__ void store_path(const WfBuf& b, uint32_t slot, const PathRec& r) {
	b.rec[slot].pA = make_float4(r.ps.ray.o.x, r.ps.ray.o.y, r.ps.ray.o.z, r.ps.ray.d.x);
	b.rec[slot].pB = make_float4(r.ps.ray.d.y, r.ps.ray.d.z, r.ps.T.x, r.ps.T.y);
	b.rec[slot].pC = make_float4(r.ps.T.z, __uint_as_float(r.pixel), __uint_as_float(r.sample), __uint_as_float(r.dim));
	b.rec[slot].pD = make_uint4(uint32_t(r.ps.bounce) | (uint32_t(r.ps.guard) << 16), r.ps.nee, __float_as_uint(r.tHit), 0u);
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
		// (block-uniform trip counts with a predicated body: every warp reaches the barrier below converged)
		uint32_t n;
		src = reinterpret_cast<const uint32_t*>(g.inst); dst = reinterpret_cast<uint32_t*>(sh.inst); n = g.n_inst * (sizeof(DInstance) / 4);
		for (uint32_t k0 = 0; k0 < n; k0 += blockDim.x) if (k0 + threadIdx.x < n) dst[k0 + threadIdx.x] = src[k0 + threadIdx.x];
		src = reinterpret_cast<const uint32_t*>(g.mat); dst = reinterpret_cast<uint32_t*>(sh.mat); n = g.n_mat * (sizeof(DMaterial) / 4);
		for (uint32_t k0 = 0; k0 < n; k0 += blockDim.x) if (k0 + threadIdx.x < n) dst[k0 + threadIdx.x] = src[k0 + threadIdx.x];
		src = reinterpret_cast<const uint32_t*>(g.vol); dst = reinterpret_cast<uint32_t*>(sh.vol); n = g.n_vol * (sizeof(DVolume) / 4);
		for (uint32_t k0 = 0; k0 < n; k0 += blockDim.x) if (k0 + threadIdx.x < n) dst[k0 + threadIdx.x] = src[k0 + threadIdx.x];
		block_sync();
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
// FAST (k_wf_scatter): medium shading through the hardware-approximation versions (ne_device.cuh "FAST medium shading"), and
// the request weight formed with two reciprocals instead of six IEEE divisions.
// INLINE (k_wf_scatter<FUSE> in mesh-free scenes, where intersectScene is a handful of analytic tests): the two next-event
// queries are answered where they arise instead of being queued for k_wf_shadow / k_wf_trfind - visibilityTr outright (no
// request, no atomic), intersectTr as far as finding the medium: the transmittance request is written with its instance,
// entry point and segment length (what k_wf_trfind would have filled in), or not at all when there is no medium to walk.
template <bool FAST, bool INLINE = false>
struct QueueSink {
	static constexpr bool kFast = FAST;
	V3 scale;
	float sel_pdf;
	const WfBuf* b;
	float* accum;
	uint32_t pixel, sample, par;
	__device__ __forceinline__ void begin() {}
	__device__ __forceinline__ void emit(V3 v) { splat(accum, pixel, v); }
	__device__ __forceinline__ V3 end(float) { return V3(0.0f); }
	__device__ __forceinline__ void light_term(const DScene& s, V3 p, V3 C, V3 f, V3 Li, float weight, float pdf, PhiloxRng&, Stats& st) {
		V3 w = FAST ? scale * (f * Li * (weight * rcp_fast(pdf) * rcp_fast(sel_pdf))) : scale * ((f * Li * weight / pdf) / sel_pdf);
		if (is_black(w)) return;
		if (INLINE) {  // visibilityTr :34-72, as k_wf_shadow runs it (visibility_tr<.., FAITHFUL = false>)
			Ray ray;
			ray.o = p;
			ray.d = C - p;
			Hit h;
			st.shadow_rays++;
			bool visible = true;
			if (intersect_scene_nomesh(s, ray, h, float(NE_EPSILON3), INFINITY, st)) {
				const int mi = s.inst[h.inst].material;
				visible = mi >= 0 && s.mat[mi].has_light;
			}
			if (visible) splat(accum, pixel, w);
			return;
		}
		uint32_t i = warp_push(&b->c->shadow);
		if (i >= b->shadowCap) { b->c->overflow = 1u; return; }
		b->sA[i] = make_float4(p.x, p.y, p.z, C.x);
		b->sB[i] = make_float4(C.y, C.z, w.x, w.y);
		b->sC[i] = make_float2(w.z, __uint_as_float(pixel));
	}
	__device__ __forceinline__ void bsdf_term(const DScene& s, Ray ray, V3 f, V3 Li, float weight, float pdf, PhiloxRng&, uint32_t stream, Stats& st) {
		if (!s.has_medium) return;  // intersectTr can only succeed through a medium (Q12)
		V3 w = FAST ? scale * (f * Li * (weight * rcp_fast(pdf) * rcp_fast(sel_pdf))) : scale * ((f * Li * weight / pdf) / sel_pdf);
		if (is_black(w)) return;
		int inst = -1;  // -1: k_wf_trfind has yet to look for the medium
		float tRemain = 0.0f;
		if (INLINE) {  // intersectTr :13-31 up to the medium, as k_wf_trfind runs it
			inst = -2;
			for (int seg = 0; seg < NE_MAX_TR_SEGMENTS; seg++) {
				Hit hh;
				st.shadow_rays++;
				if (!intersect_scene_nomesh(s, ray, hh, float(NE_EPSILON3), INFINITY, st)) break;
				const int mi = s.inst[hh.inst].material;
				if (mi >= 0 && s.mat[mi].has_medium && s.mat[mi].volume >= 0) {
					inst = hh.inst;
					ray.o = ray.at(hh.tNear);
					tRemain = hh.tFar - hh.tNear;
					break;
				}
				if (mi >= 0 && s.mat[mi].has_medium) {  // HomogeneousMedia: closed-form transmittance, nothing to walk
					splat(accum, pixel, w * homog_tr(s.mat[mi], hh.tFar - hh.tNear));
					break;
				}
				ray.o = hh.p;
			}
			if (inst < 0) return;  // no grid medium along the ray: nothing for k_wf_tr to do
		}
		uint32_t i = warp_push(&b->c->tr);
		if (i >= b->trCap) { b->c->overflow = 1u; return; }
		b->tA[par][i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
		b->tB[par][i] = make_float4(ray.d.y, ray.d.z, w.x, w.y);
		b->tC[par][i] = make_float4(w.z, __uint_as_float(pixel), __uint_as_float(sample), __uint_as_float(stream));
		b->tD[par][i] = make_float4(1.0f, tRemain, __int_as_float(inst), __uint_as_float(0u));
	}
};

// ---------------------------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// One thread: reset the bookkeeping and take this render's arguments. The pool itself is not touched: slots are handed
// out from a bump counter until the first ones come back through the free stack.
__global__ void k_wf_init(WfBuf b, unsigned long long workTotal, WfDyn dyn) {
	WfCounts c;
	memset(&c, 0, sizeof(c));
	c.workTotal = workTotal;
	c.dyn = dyn;
	c.lastNs = global_ns();
	*b.c = c;
}

// Per-stage device time without extra launches: the kernels of an iteration run back to back on one stream, so the first
// thread of each, before anything else, books the time since the previous stamp on the account of the stage that just
// ended (the kind of the kernel launched before it; the host names it in WfParams::prevStage).
__device__ __forceinline__ void stage_stamp(const WfBuf& b, int prevStage) {
	if (prevStage >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
		WfCounts& c = *b.c;
		unsigned long long now = global_ns();
		c.stageNs[prevStage] += now - c.lastNs;
		c.lastNs = now;
	}
}

// One thread: retire the finished iteration (flip the ping-pong pairs: what was "next" is current now; clear the stage
// queues) and plan the refill of the pool. Inside the render graph it also decides whether the WHILE node runs its body
// again (cudaGraphSetConditional); the host-driven loop reads the mapped `hostDone` word instead.
__global__ void k_wf_plan(WfBuf b, cudaGraphConditionalHandle loop, int inGraph, volatile uint32_t* hostDone, uint32_t launchesPerIteration, int prevStage) {
	stage_stamp(b, prevStage);
	WfCounts& c = *b.c;
	c.par ^= 1u;
	c.extend = c.next;
	c.next = 0;
	c.vol = c.volNext;
	c.volNext = 0;
	c.tr = min(c.trNext, b.carryCap);  // reservations beyond the room for carried walks were never written (k_wf_tr)
	c.trNew0 = c.tr;  // requests pushed from here on are new: k_wf_trfind locates their medium
	c.trNext = 0;
	c.scat = c.surf = c.shadow = 0;
	c.volHead = c.trHead = c.extHead = c.shHead = c.trfHead = 0;
	const unsigned long long remaining = c.workTotal - c.workNext;
	const unsigned long long room = (unsigned long long)c.freeN + (b.nSlots - c.bump);
	const uint32_t gen = uint32_t(remaining < room ? remaining : room);
	c.gen = gen;
	c.freeTake = gen < c.freeN ? gen : c.freeN;
	c.freeN -= c.freeTake;
	c.bumpBase = c.bump;
	c.bump += gen - c.freeTake;
	c.done = (c.extend == 0 && gen == 0 && c.vol == 0 && c.tr == 0) ? 1u : 0u;
	if (!c.done) {
		c.iterations++;
		c.launches += launchesPerIteration;
	}
	if (hostDone) *hostDone = c.done;
	if (inGraph) cudaGraphSetConditional(loop, c.done ? 0u : 1u);
}

// One thread, after the loop: fold this render's bookkeeping into the context's counters.
__global__ void k_wf_finish(WfBuf b, DCounters* counters, int foldTimes) {
	WfCounts& c = *b.c;
	unsigned long long now = global_ns();
	c.stageNs[STAGE_OTHER] += now - c.lastNs;
	if (foldTimes)
		for (int k = 0; k < STAGE_KINDS; k++) counters->stage_ns[k] += c.stageNs[k];
	counters->iterations += c.iterations;
	counters->launches += c.launches + 3;  // + init, the first plan, this kernel
	if (c.overflow) counters->overflow = 1;
}

// OfflineEngine.cpp:64-67 + Li's first intersectScene (:187-193, :244-260) fused: sample jitter,
// Camera::getRayPassingThrough, trace and classify `gen` new camera paths. A camera path that ends right there (it
// misses everything, or sees an emitter) never touches the pool - no slot, no record, no queue entry; only survivors
// take a slot (from the `gen` the plan set aside; commit returns the rest) and are written ONCE, path and hit
// together, straight into the volume / surface queue. In the C2 frame 5 of 6 camera paths miss the medium's box.
template <int MINB>  // resident blocks per SM: 3 without meshes, 4 with (the BVH walk is latency-bound: warps in flight count)
__global__ void __launch_bounds__(256, MINB) k_wf_generate(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	NE_STAGE_SCENE();
	const uint32_t gen = b.c->gen;
	if (gen == 0) return;
	const uint32_t freeN = b.c->freeN, freeTake = b.c->freeTake, bumpBase = b.c->bumpBase, par = b.c->par;
	const uint32_t vol0 = b.c->vol;  // (read only where it is constant for the whole kernel: P.genToVol)
	const unsigned long long workBase = b.c->workNext;
	// the camera of this render, staged in shared memory (19 floats that would otherwise sit in registers for the whole kernel)
	__shared__ DCamera cam;
	if (threadIdx.x < sizeof(DCamera) / 4) reinterpret_cast<uint32_t*>(&cam)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&b.c->dyn.cam)[threadIdx.x];
	block_sync();
	float* const accum = b.c->dyn.accum;
	const unsigned long long seed = b.c->dyn.seed;
	const uint32_t sppBegin = uint32_t(b.c->dyn.sppBegin);
	// work items enumerate (sample, pixel of the culling rectangle); the rectangle is the whole frame unless cull_rect found less
	const uint32_t rx0 = uint32_t(b.c->dyn.rx0), ry0 = uint32_t(b.c->dyn.ry0), rw = uint32_t(b.c->dyn.rw), rh = uint32_t(b.c->dyn.rh);
	const uint32_t npix = rw * rh;
	const bool tiled = (rw % 8 == 0) && (rh % 4 == 0);  // else the rectangle is walked row by row (any bijection will do: Philox is keyed by pixel)
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < gen; i += gridDim.x * blockDim.x) {
		unsigned long long w = workBase + i;
		const uint32_t p = uint32_t(w % npix);
		uint32_t px = p % rw, py = p / rw;
		if (tiled) {  // a warp's 32 consecutive work items cover an 8x4 pixel tile, not a 32x1 strip: coherent camera rays
			const uint32_t t = p >> 5, l = p & 31u, tilesX = rw >> 3;
			px = (t % tilesX) * 8u + (l & 7u);
			py = (t / tilesX) * 4u + (l >> 3);
		}
		const int x = int(rx0 + px), y = int(ry0 + py);
		const uint32_t pixel = uint32_t(y) * uint32_t(P.W) + uint32_t(x);
		uint32_t sample = sppBegin + uint32_t(w / npix);
		PhiloxRng rng;
		rng.init(seed, pixel, sample);
		float u = float(float(x) + rng.next()) / float(P.W);
		float v = float(float(y) + rng.next()) / float(P.H);
		PathRec r;
		r.ps.ray = camera_ray(cam, u, v, rng);
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
		QueueSink<false> sink;
		sink.accum = accum;
		sink.pixel = pixel;
		int kind = classify_hit(S, did, h, r.ps, sink);
		if (kind == HIT_TERMINATE) continue;
		// the j-th survivor takes the j-th reserved slot: first the freeTake entries popped off the free stack, then fresh ones
		const uint32_t j = warp_push(&b.c->genTaken);
		const uint32_t slot = j < freeTake ? b.qFree[freeN + freeTake - 1u - j] : bumpBase + (j - freeTake);
		store_path(b, slot, r);
		store_hit(b, slot, h);
		if (P.genToVol) {
			// the j-th survivor is also the j-th new entry of the volume queue (nobody else appends to it during this kernel;
			// k_wf_commit adds genTaken to the count): one same-address atomic per warp instead of two (41 % of this kernel's
			// stall samples sat on them)
			b.qVol[par][vol0 + j] = slot;
		} else if (kind == HIT_VOLUME) {
			if (S.mat[S.inst[h.inst].material].volume >= 0) b.qVol[par][warp_push(&b.c->vol)] = slot;
			else b.qScat[warp_push(&b.c->scat)] = slot;
		} else b.qSurf[warp_push(&b.c->surf)] = slot;
	}
	flush_stats_wf(st, P.counters);
}
// One thread: publish the refill (after generate has read the old counts); reserved slots no survivor took go back.
__global__ void k_wf_commit(WfBuf b, DCounters* counters, int prevStage, int genToVol) {
	stage_stamp(b, prevStage);
	WfCounts& c = *b.c;
	if (genToVol) c.vol += c.genTaken;  // k_wf_generate wrote the survivors' queue entries itself
	if (c.genTaken <= c.freeTake) {  // the untouched part of the free-stack reservation is still in place; no fresh slot was used
		c.freeN += c.freeTake - c.genTaken;
		c.bump = c.bumpBase;
	} else c.bump = c.bumpBase + (c.genTaken - c.freeTake);
	c.workNext += c.gen;
	if (c.gen) atomicAdd(&counters->paths, (unsigned long long)c.gen);
	c.gen = 0;
	c.genTaken = 0;
	c.freeTake = 0;
}

// Scene::intersectScene for every path of the extend queue + classify (Li :187-193, :244-260).
__global__ void __launch_bounds__(256) k_wf_extend(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	NE_STAGE_SCENE();
	const uint32_t n = b.c->extend;
	if (n == 0) return;
	const uint32_t par = b.c->par;
	float* const accum = b.c->dyn.accum;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qExt[par][i];
		if (slot == NE_Q_INVALID) continue;  // the unused tail of a tracking warp's last chunk (WarpChunk)
		PathRec r = load_path(b, slot);
		Hit h;
		st.extend_rays++;
		bool did = intersect_scene(S, r.ps.ray, h, float(NE_EPSILON12), INFINITY, st);
		QueueSink<false> sink;
		sink.accum = accum;
		sink.pixel = r.pixel;
		int kind = classify_hit(S, did, h, r.ps, sink);
		if (kind == HIT_TERMINATE) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			store_hit(b, slot, h);
			if (kind == HIT_VOLUME) {
				// a grid medium is walked by k_wf_track; a HomogeneousMedia needs no walk and goes straight to k_wf_scatter
				if (S.mat[S.inst[h.inst].material].volume >= 0) b.qVol[par][warp_push(&b.c->vol)] = slot;
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

// The persistent kernels' hot queues, reserved in CHUNKS per warp. A refill used to cost one global atomic per queue for a
// handful of entries (8 on average): with 4736 warps at it, the same-address atomics of volHead / scat / next queued up in L2
// (18.7 % of k_wf_track's stall samples on its incoherent launches). A warp now reserves NE_QCHUNK entries at a time and hands
// them to its lanes from registers; what is left of a chunk is used before the next one is touched. For an OUTPUT queue the
// unused tail of a warp's last chunk is filled with NE_Q_INVALID at the end of the kernel (finish()): consumers skip such
// entries (at most NE_QCHUNK - 1 per warp and queue, at the end of a chunk: a few partially filled consumer warps).
struct WarpChunk {
	uint32_t base = 0, left = 0;  // warp-uniform
	// entries for the lanes in `mine`; all 32 lanes call it. Returns this lane's entry (meaningful for lanes in the mask).
	__device__ __forceinline__ uint32_t take(uint32_t* counter, bool mine) {
		const unsigned mask = __ballot_sync(0xffffffffu, mine);
		const uint32_t m = __popc(mask);
		if (m == 0) return 0;
		const uint32_t rank = __popc(mask & ((1u << (threadIdx.x & 31)) - 1));
		uint32_t idx = base + rank;  // from what is left of the current chunk
		if (m > left) {
			uint32_t nb = 0;
			if ((threadIdx.x & 31) == 0) nb = atomicAdd(counter, NE_QCHUNK);
			nb = __shfl_sync(0xffffffffu, nb, 0);
			if (rank >= left) idx = nb + (rank - left);
			base = nb + (m - left);
			left = NE_QCHUNK - (m - left);
		} else {
			base += m;
			left -= m;
		}
		return idx;
	}
	// output queues: the rest of the last chunk holds no slot
	__device__ __forceinline__ void finish(uint32_t* queue) {
		for (uint32_t i = threadIdx.x & 31; i < left; i += 32) queue[base + i] = NE_Q_INVALID;
		left = 0;
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
// TRACK_*_SM variants: every block copies the 2-byte majorant tables of the scene's volumes (64 KB for a 256^3 grid) into
// its shared memory with bulk async copies (cp.async.bulk, completion counted by an mbarrier) and points the staged
// DVolume entries at the copy, so the one load on a walk's critical path - the majorant at every brick crossing - is a
// shared-memory access instead of an L1/L2 round trip (18.5 % of k_wf_track's stall samples in round 1). The host picks
// these variants when the tables fit (wavefront_render); the kernels then run ONE 1024-thread block per SM, which is
// the same 32 warps and 64 registers per thread as four 256-thread blocks, with one copy of the table instead of four.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t maj_table_bytes(const DVolume& v) { return (uint32_t((v.bx + 2) * (v.by + 2) * (v.bz + 2)) * 2u + 15u) & ~15u; }
__device__ __forceinline__ void stage_majorants(const DScene& g, SceneCache& sh, unsigned char* table, unsigned long long* bar) {
	const uint32_t barS = smem_u32(bar);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barS), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	block_sync();
	if (threadIdx.x == 0) {
		uint32_t total = 0;
		for (int i = 0; i < g.n_vol; i++) total += maj_table_bytes(g.vol[i]);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barS), "r"(total) : "memory");
		uint32_t off = 0;
		for (int i = 0; i < g.n_vol; i++) {
			const uint32_t bytes = maj_table_bytes(g.vol[i]);
			const unsigned char* src = reinterpret_cast<const unsigned char*>(g.vol[i].maj16);
			for (uint32_t o = 0; o < bytes; o += 32768u) {
				const uint32_t n = min(32768u, bytes - o);
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(table + off + o)),
				             "l"(src + o), "r"(n), "r"(barS)
				             : "memory");
			}
			sh.vol[i].maj16 = reinterpret_cast<const unsigned short*>(table + off);
			off += bytes;
		}
	}
	block_sync();  // the patched table pointers
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"NE_MAJ_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra NE_MAJ_DONE;\n"
		"bra NE_MAJ_WAIT;\n"
		"NE_MAJ_DONE:\n"
		"}\n" ::"r"(barS),
		"r"(0)
		: "memory");
}
#define NE_STAGE_MAJORANTS(MODE)                                              \
	extern __shared__ __align__(128) unsigned char majTable_[];              \
	__shared__ unsigned long long majBar_;                                    \
	if (NE_TRACK_IS_SM(MODE)) stage_majorants(P.s, sceneCache_, majTable_, &majBar_)

template <int BRICKMAJ, int THREADS, int MINB>  // TRACK_GLOBAL / TRACK_BRICK / TRACK_SKIP / TRACK_*_SM (ne_tracking.cuh)
__global__ void __launch_bounds__(THREADS, MINB) k_wf_track(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	const uint32_t n = b.c->vol;
	if (n == 0) return;
	NE_STAGE_SCENE();
	NE_STAGE_MAJORANTS(BRICKMAJ);
	typedef typename WalkRngOf<(BRICKMAJ != 0), PhiloxRng>::type WalkRng;
	const uint32_t par = b.c->par;
	const unsigned long long seed = b.c->dyn.seed;
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
	WarpChunk cScat, cNext, cFetch;  // the three queues every refill touches, reserved in chunks (WarpChunk)
	while (true) {
		unsigned walking = __ballot_sync(0xffffffffu, state == L_MOVING || state == L_CAND);
		if (walking == 0 || (!exhausted && 32 - __popc(walking) >= P.refill)) {
			// ---- finish + refill
			const bool fin = state >= L_FIN_HIT;
			const bool esc = state == L_FIN_END;
			const uint32_t bg2 = bg + 65536u;  // volume_escape (Li :209-213, Q1): the guard lives in bits 16..31
			const bool dead = esc && (bg2 >> 16) > NE_MAX_NULL_SEGMENTS;
			WarpReserve rVolNext, rFree;  // rare: walks cut in the tail, paths that ran out of null segments
			rVolNext.issue(&b.c->volNext, state == L_FIN_BUDGET);  // (at most one entry per live slot: cannot overflow)
			rFree.issue(&b.c->freeN, dead);
			// the shuffles are warp-wide: resolve every reservation before the lanes part ways
			const uint32_t iScat = cScat.take(&b.c->scat, state == L_FIN_HIT), iNext = cNext.take(&b.c->next, esc && !dead);
			const uint32_t iFetch = cFetch.take(&b.c->volHead, !exhausted && (state == L_IDLE || fin));
			const uint32_t iVolNext = rVolNext.get(), iFree = rFree.get();
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
					b.qVol[par ^ 1u][iVolNext] = slot;
				} else if (dead) {
					b.qFree[iFree] = slot;
				} else {
					V3 o = ray.at(tFar + 0.01f);  // step past the far side, same bounce
					b.rec[slot].pA = make_float4(o.x, o.y, o.z, ray.d.x);
					b.rec[slot].pD.x = bg2;
					b.qExt[par ^ 1u][iNext] = slot;
				}
				b.rec[slot].pC.w = __uint_as_float(rng.dim);
				b.rec[slot].hA.w = 0.0f;
				state = L_IDLE;
			}
			if (!exhausted) {
				const uint32_t i = iFetch;
				if (state == L_IDLE && i < n) {
					slot = b.qVol[par][i];
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
					rng.init(seed, __float_as_uint(C.y), __float_as_uint(C.z), __float_as_uint(C.w));
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
				if (budget-- <= 0 && (exhausted || P.cutAlways)) state = L_FIN_BUDGET;
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
	cScat.finish(b.qScat);
	cNext.finish(b.qExt[par ^ 1u]);
	flush_stats_wf(st, P.counters);
}

// Real collisions: phase function, next-event setup, continuation (Li :215-236).
// FUSE (scenes without meshes, where intersectScene is a handful of analytic tests): the scattered ray is traced and
// classified right here, so a path that goes on through a grid medium enters the NEXT iteration's volume queue
// directly, one that reaches a surface enters this iteration's surface queue, and one that ends frees its slot: no
// trip through the extend queue, i.e. one 128-byte record read and one hit write fewer per scatter event.
// LS: the scene's light set (-1 = any; else every light is a DiffuseLight on that primitive kind: the other kinds, the
// directional / environment code and - with them - the HomogeneousMedia branch are compiled out; ne_device.cuh light_sample_point)
// FASTSH: FAST medium shading (ne_device.cuh), the default; NE_B200_EXACT_SHADING=1 runs the reference-order arithmetic.
#ifndef NE_SCATTER_BLOCKS
#define NE_SCATTER_BLOCKS 2  // resident blocks per SM k_wf_scatter is compiled for
#endif
template <bool FUSE, int LS, bool FASTSH>
__global__ void __launch_bounds__(256, NE_SCATTER_BLOCKS) k_wf_scatter(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	NE_STAGE_SCENE();
	const uint32_t n = b.c->scat;
	if (n == 0) return;
	const uint32_t par = b.c->par;
	float* const accum = b.c->dyn.accum;
	const unsigned long long seed = b.c->dyn.seed;
	const int bounces = b.c->dyn.bounces;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qScat[i];
		if (slot == NE_Q_INVALID) continue;  // the unused tail of a tracking warp's last chunk (WarpChunk)
		PathRec r = load_path(b, slot);
		Hit h = load_hit(b, slot);
		PhiloxRng rng;
		rng.init(seed, r.pixel, r.sample, r.dim);
		QueueSink<FASTSH, FUSE> sink;
		sink.b = &b;
		sink.accum = accum;
		sink.pixel = r.pixel;
		sink.sample = r.sample;
		sink.par = par;
		int next;
		if (LS < 0 && S.mat[S.inst[h.inst].material].volume < 0) {  // the specialised variants run only in scenes without HomogeneousMedia
			next = shade_volume_homog(S, r.ps, h, rng, sink, st);
		} else {
			Ray rayO = transform_ray(r.ps.ray, S.inst[h.inst].Mi);
			next = volume_scatter<LS>(S, r.ps, h, rayO, r.tHit, rng, sink, st);
		}
		if (next == PATH_NEXT_BOUNCE) r.ps.bounce++;
		if (next == PATH_DONE || r.ps.bounce >= bounces) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			r.dim = rng.dim;
			if (FUSE) {
				Hit h2;
				const uint32_t prims0 = st.prim_tests;
				bool did = intersect_scene_nomesh(S, r.ps.ray, h2, float(NE_EPSILON12), INFINITY, st);
				int kind = HIT_TERMINATE;
				const bool grid = did && h2.inst >= 0 && S.inst[h2.inst].material >= 0 && S.mat[S.inst[h2.inst].material].has_bsdf &&
				                  S.mat[S.inst[h2.inst].material].transmissive && S.mat[S.inst[h2.inst].material].volume >= 0;
				// only a path that goes on through a grid medium (or ends) is settled here. A HomogeneousMedia hit belongs in
				// the scatter queue this kernel is draining, and a surface hit would be shaded a second time in this iteration
				// (the request arrays are sized for one shading event per live slot): both take the trip through k_wf_extend
				const bool ends = !did || is_black(r.ps.T) || (!grid && (S.inst[h2.inst].material < 0 || !S.mat[S.inst[h2.inst].material].has_bsdf));
				if (grid || ends) kind = classify_hit(S, did, h2, r.ps, sink);
				else {
					st.prim_tests = prims0;  // k_wf_extend will count the query
					store_path(b, slot, r);
					b.qExt[par ^ 1u][warp_push(&b.c->next)] = slot;
					continue;
				}
				st.extend_rays++;
				if (kind == HIT_TERMINATE) {
					b.qFree[warp_push(&b.c->freeN)] = slot;
					continue;
				}
				store_path(b, slot, r);
				store_hit(b, slot, h2);
				b.qVol[par ^ 1u][warp_push(&b.c->volNext)] = slot;
				continue;
			}
			store_path(b, slot, r);
			b.qExt[par ^ 1u][warp_push(&b.c->next)] = slot;
		}
	}
	flush_stats_wf(st, P.counters);
}

// Surface hits: GGX shading, next-event setup, continuation (Li :262-283).
template <int LS>  // the scene's light set, as for k_wf_scatter
__global__ void __launch_bounds__(256, 3) k_wf_surface(WfBuf b, WfParams P) {  // 3 resident blocks (80 registers), as measured best for the GGX shading chain
	stage_stamp(b, P.prevStage);
	NE_STAGE_SCENE();
	const uint32_t n = b.c->surf;
	if (n == 0) return;
	const uint32_t par = b.c->par;
	float* const accum = b.c->dyn.accum;
	const unsigned long long seed = b.c->dyn.seed;
	const int bounces = b.c->dyn.bounces;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qSurf[i];
		PathRec r = load_path(b, slot);
		Hit h = load_hit(b, slot);
		PhiloxRng rng;
		rng.init(seed, r.pixel, r.sample, r.dim);
		QueueSink<false> sink;
		sink.b = &b;
		sink.accum = accum;
		sink.pixel = r.pixel;
		sink.sample = r.sample;
		sink.par = par;
		int next = shade_surface<PhiloxRng, LS>(S, r.ps, h, rng, sink, st);
		if (next == PATH_NEXT_BOUNCE) r.ps.bounce++;
		if (next == PATH_DONE || r.ps.bounce >= bounces) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			r.dim = rng.dim;
			store_path(b, slot, r);
			b.qExt[par ^ 1u][warp_push(&b.c->next)] = slot;
		}
	}
	flush_stats_wf(st, P.counters);
}

// visibilityTr requests: splat the weight when nothing or an emitter is hit first.
__global__ void __launch_bounds__(256) k_wf_shadow(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	NE_STAGE_SCENE();
	const uint32_t n = min(b.c->shadow, b.shadowCap);
	if (n == 0) return;
	float* const accum = b.c->dyn.accum;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 A = b.sA[i], B = b.sB[i];
		float2 C = b.sC[i];
		PhiloxRng dummy;
		float vis = visibility_tr<PhiloxRng, false, true>(S, V3(A.x, A.y, A.z), V3(A.w, B.x, B.y), dummy, st);
		if (vis != 0) splat(accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * vis);
	}
	flush_stats_wf(st, P.counters);
}

// intersectTr :13-31 for the requests pushed in this iteration: march THROUGH non-medium surfaces until a medium
// (the request gets its instance, entry point and segment length) or nothing (the request is dropped: Li = 0).
__global__ void __launch_bounds__(256) k_wf_trfind(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	NE_STAGE_SCENE();
	const uint32_t first = b.c->trNew0, n = min(b.c->tr, b.trCap);
	if (n <= first) return;
	const uint32_t par = b.c->par;
	float* const accum = b.c->dyn.accum;
	Stats st;
	st.clear();
	for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if (__float_as_int(b.tD[par][i].z) != -1) continue;  // k_wf_scatter<FUSE> found the medium itself (QueueSink<.., INLINE>)
		float4 A = b.tA[par][i], B = b.tB[par][i];
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
				float4 C = b.tC[par][i];
				splat(accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * homog_tr(S.mat[mi], hh.tFar - hh.tNear));
				break;  // inst stays -2: the request is done
			}
			ray.o = hh.p;
		}
		b.tA[par][i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
		b.tD[par][i] = make_float4(1.0f, tRemain, __int_as_float(inst), __uint_as_float(0u));
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
	__device__ __forceinline__ bool fetch(const WfBuf& b, uint32_t i, SceneTrace& tr, Stats& st) {
		slot = b.qExt[b.c->par][i];
		if (slot == NE_Q_INVALID) return false;  // the unused tail of a tracking warp's last chunk (WarpChunk)
		float4 A = b.rec[slot].pA, B = b.rec[slot].pB, C = b.rec[slot].pC;
		bounce = int(b.rec[slot].pD.x & 0xffffu);
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y);
		T = V3(B.z, B.w, C.x);
		pixel = __float_as_uint(C.y);
		st.extend_rays++;
		tr.begin(ray, float(NE_EPSILON12), INFINITY);
		return true;
	}
	__device__ __forceinline__ bool consume(const DScene& S, const WfBuf& b, const WfParams& P, SceneTrace& tr, Stats&) {
		PathState ps;
		ps.ray = tr.rayW;
		ps.T = T;
		ps.bounce = bounce;
		QueueSink<false> sink;
		sink.accum = b.c->dyn.accum;
		sink.pixel = pixel;
		int kind = classify_hit(S, tr.did, tr.hit, ps, sink);
		if (kind == HIT_TERMINATE) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			store_hit(b, slot, tr.hit);
			if (kind == HIT_VOLUME) {
				if (S.mat[S.inst[tr.hit.inst].material].volume >= 0) b.qVol[b.c->par][warp_push(&b.c->vol)] = slot;
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
	__device__ __forceinline__ static uint32_t count(const WfBuf& b) { return min(b.c->shadow, b.shadowCap); }
	__device__ __forceinline__ bool fetch(const WfBuf& b, uint32_t i, SceneTrace& tr, Stats& st) {
		req = i;
		float4 A = b.sA[i], B = b.sB[i];
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y) - ray.o;
		st.shadow_rays++;
		tr.begin(ray, float(NE_EPSILON3), INFINITY);
		return true;
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
			splat(b.c->dyn.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x));
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
	__device__ __forceinline__ static uint32_t count(const WfBuf& b) { return min(b.c->tr, b.trCap) - b.c->trNew0; }
	__device__ __forceinline__ bool fetch(const WfBuf& b, uint32_t i, SceneTrace& tr, Stats& st) {
		req = i;
		seg = 0;
		const uint32_t par = b.c->par;
		if (__float_as_int(b.tD[par][i].z) != -1) return false;  // k_wf_scatter<FUSE> found the medium itself (QueueSink<.., INLINE>)
		float4 A = b.tA[par][i], B = b.tB[par][i];
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y);
		st.shadow_rays++;
		tr.begin(ray, float(NE_EPSILON3), INFINITY);
		return true;
	}
	__device__ __forceinline__ bool consume(const DScene& S, const WfBuf& b, const WfParams& P, SceneTrace& tr, Stats& st) {
		Ray ray = tr.rayW;
		int inst = -2;
		float tRemain = 0;
		const uint32_t par = b.c->par;
		if (tr.did) {
			const Hit& hh = tr.hit;
			int mi = S.inst[hh.inst].material;
			if (mi >= 0 && S.mat[mi].has_medium && S.mat[mi].volume >= 0) {
				inst = hh.inst;
				ray.o = ray.at(hh.tNear);
				tRemain = hh.tFar - hh.tNear;
			} else if (mi >= 0 && S.mat[mi].has_medium) {  // HomogeneousMedia: closed-form transmittance, nothing to walk
				float4 B = b.tB[par][req], C = b.tC[par][req];
				splat(b.c->dyn.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * homog_tr(S.mat[mi], hh.tFar - hh.tNear));
			} else {
				ray.o = hh.p;
				if (++seg < NE_MAX_TR_SEGMENTS) {  // through the surface: next segment of the same request
					st.shadow_rays++;
					tr.begin(ray, float(NE_EPSILON3), INFINITY);
					return true;
				}
			}
		}
		b.tA[par][req] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
		b.tD[par][req] = make_float4(1.0f, tRemain, __int_as_float(inst), __uint_as_float(0u));
		return false;
	}
};

#ifndef NE_TRACE_BLOCKS
#define NE_TRACE_BLOCKS 4  // resident 256-thread blocks per SM the trace kernels are compiled for (64 registers: the walk is bound by
                           // the latency of dependent node loads, so warps in flight count; measured 2: 43.5, 3: 33.9, 4: 31.0, 5: 33.3 ms on C4)
#endif
template <class JOB>
__global__ void __launch_bounds__(256, NE_TRACE_BLOCKS) k_wf_trace(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	const uint32_t first = JOB::first(b), n = JOB::count(b);
	if (n == 0) return;
	NE_STAGE_SCENE();
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
					const bool wanted = state == T_IDLE;
					if (wanted && i < n && job.fetch(b, first + i, tr, st)) state = T_FOLD;  // (an entry may hold no slot: the lane asks again)
					if (__ballot_sync(0xffffffffu, wanted && i >= n)) exhausted = true;  // a lane came back empty-handed
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
template <int BRICKMAJ, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_wf_tr(WfBuf b, WfParams P) {
	stage_stamp(b, P.prevStage);
	const uint32_t n = min(b.c->tr, b.trCap);
	if (n == 0) return;
	NE_STAGE_SCENE();
	NE_STAGE_MAJORANTS(BRICKMAJ);
	typedef typename WalkRngOf<(BRICKMAJ != 0), PhiloxRng>::type WalkRng;
	const uint32_t par = b.c->par;
	const unsigned long long seed = b.c->dyn.seed;
	float* const accum = b.c->dyn.accum;
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
	WarpChunk cFetch;  // request indices, reserved in chunks
	while (true) {
		unsigned walking = __ballot_sync(0xffffffffu, state == L_MOVING || state == L_CAND);
		if (walking == 0 || (!exhausted && 32 - __popc(walking) >= P.refill)) {
			// ---- finish + refill
			// walks are cut only in the kernel's tail (or always, in test mode): no atomic is issued here otherwise
			WarpReserve rNext;
			rNext.issue(&b.c->trNext, state == L_FIN_BUDGET);
			const uint32_t iNext = rNext.get();
			if (state == L_FIN_BUDGET && iNext >= b.carryCap) {
				// no room to carry the walk over: it simply goes on where it stands (the budget is a scheduling device)
				budget = P.budget;
				state = L_MOVING;
			}
			const bool fin = state >= L_FIN_HIT;
			const uint32_t iFetch = cFetch.take(&b.c->trHead, !exhausted && (state == L_IDLE || fin));  // warp-wide: before the lanes part ways
			if (fin) {
				float4 B = b.tB[par][req], C = b.tC[par][req];
				if (state == L_FIN_BUDGET) {
					uint32_t j = iNext;
					V3 o = ray.at(trk.t);
					b.tA[par ^ 1u][j] = make_float4(o.x, o.y, o.z, ray.d.x);
					b.tB[par ^ 1u][j] = B;
					b.tC[par ^ 1u][j] = C;
					b.tD[par ^ 1u][j] = make_float4(Tr, tRemain - trk.t, __int_as_float(inst), __uint_as_float(rng.dim));
				} else if (Tr != 0) {
					splat(accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * V3(Tr));
				}
				state = L_IDLE;
			}
			if (!exhausted) {
				const uint32_t i = iFetch;
				if (state == L_IDLE && i < n) {
					float4 D = b.tD[par][i];
					inst = __float_as_int(D.z);
					if (inst >= 0) {  // else no medium along the ray: the request is dropped
						float4 A = b.tA[par][i], B = b.tB[par][i], C = b.tC[par][i];
						req = i;
						ray.o = V3(A.x, A.y, A.z);
						ray.d = V3(A.w, B.x, B.y);
						Tr = D.x;
						tRemain = D.y;
						rng.init(seed, __float_as_uint(C.y), __float_as_uint(C.z), __float_as_uint(D.w), __float_as_uint(C.w));
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
				if (budget-- <= 0 && (exhausted || P.cutAlways)) state = L_FIN_BUDGET;
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

// What decides which kernels one wavefront iteration launches and with which fixed parameters: the render graph is
// rebuilt when any of it changes.
struct WfVariant {
	unsigned long long sceneGen;
	int W, H;
	int trace, fuse, brick, skip, sm, genBlocks, lightSet, stamps, media, surfaces, foldTimes, fastShade;
	int budget, refill, moves, walkBudget, walkRefill, cutAlways, l2persist, genToVol;
	uint32_t nSlots;
};

struct ne_wavefront_state {
	WfBuf b{};
	uint32_t nSlots = 0;
	std::vector<void*> allocs;
	uint32_t* hostDone = nullptr;  // pinned, mapped (host-driven loop only)
	uint32_t* devDone = nullptr;
	int gridBlocks = 0, smCount = 0;
	size_t smemOptin = 0;
	std::vector<cudaEvent_t> events;
	// the render graph: first plan -> WHILE(not done){ one wavefront iteration; plan } -> finish
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	cudaStream_t capStream = nullptr;
	WfVariant built{};
	bool haveGraph = false, graphBroken = false;
};

namespace ne {

static void graph_free(ne_wavefront_state* w) {
	if (w->exec) cudaGraphExecDestroy(w->exec);
	if (w->graph) cudaGraphDestroy(w->graph);
	w->exec = nullptr;
	w->graph = nullptr;
	w->haveGraph = false;
}

static void wavefront_free_state(ne_wavefront_state* w) {
	if (!w) return;
	graph_free(w);
	if (w->capStream) cudaStreamDestroy(w->capStream);
	for (void* p : w->allocs) cudaFree(p);
	if (w->hostDone) cudaFreeHost(w->hostDone);
	for (cudaEvent_t e : w->events) cudaEventDestroy(e);
	delete w;
}

size_t wavefront_record_bytes() { return sizeof(PathSlot); }

void wavefront_free(ne_b200_ctx* ctx) {
	for (int l = 0; l < 2; l++) {
		wavefront_free_state(ctx->wf[l]);
		ctx->wf[l] = nullptr;
		if (ctx->laneStream[l]) cudaStreamDestroy(ctx->laneStream[l]);
		if (ctx->laneJoin[l]) cudaEventDestroy(ctx->laneJoin[l]);
		ctx->laneStream[l] = nullptr;
		ctx->laneJoin[l] = nullptr;
	}
	if (ctx->laneFork) cudaEventDestroy(ctx->laneFork);
	ctx->laneFork = nullptr;
}

template <class T>
static cudaError_t wf_alloc(ne_wavefront_state* w, T** p, size_t n) {
	cudaError_t e = cudaMalloc(p, n * sizeof(T));
	if (e == cudaSuccess) w->allocs.push_back(*p);
	return e;
}

// Builds the pool for `nSlots` path slots into a fresh state object; the context only sees it once it is complete.
static cudaError_t wavefront_build(ne_b200_ctx* ctx, uint32_t nSlots, ne_wavefront_state** out) {
	ne_wavefront_state* w = new ne_wavefront_state();
	w->nSlots = nSlots;
	WfBuf& b = w->b;
	b.nSlots = nSlots;
	b.shadowCap = nSlots;
	b.carryCap = std::max(nSlots / 4, 1u << 18);
	b.trCap = nSlots + b.carryCap;
	cudaError_t e = cudaSuccess;
	// queues the tracking kernels fill in per-warp chunks (WarpChunk) hold, beside one entry per live slot, the unused tails of
	// the warps' last chunks: at most NE_QCHUNK per warp of the largest persistent grid (a few hundred thousand entries)
	const size_t slack = size_t(NE_QCHUNK) * 64 * 256;
#define A(field, n) if (e == cudaSuccess) e = wf_alloc(w, &b.field, n);
	A(rec, nSlots) A(sA, b.shadowCap) A(sB, b.shadowCap) A(sC, b.shadowCap)
	for (int k = 0; k < 2; k++) {
		A(tA[k], b.trCap) A(tB[k], b.trCap) A(tC[k], b.trCap) A(tD[k], b.trCap)
		A(qExt[k], nSlots + slack) A(qVol[k], nSlots)
	}
	A(qScat, nSlots + slack) A(qSurf, nSlots) A(qFree, nSlots) A(c, 1)
#undef A
	if (e == cudaSuccess) e = cudaHostAlloc(&w->hostDone, sizeof(uint32_t), cudaHostAllocMapped);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer(&w->devDone, w->hostDone, 0);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->capStream, cudaStreamNonBlocking);
	cudaDeviceProp prop;
	if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, ctx->device);
	if (e != cudaSuccess) {
		wavefront_free_state(w);
		return e;
	}
	w->smCount = prop.multiProcessorCount;
	w->gridBlocks = prop.multiProcessorCount * 8;  // 148 SMs x 8 resident 256-thread blocks
	w->smemOptin = prop.sharedMemPerBlockOptin;
	*out = w;
	return cudaSuccess;
}

// A pool that does not fit (a smaller or busy GPU: the default is 64 Mi slots, ~29 GB) is retried at half the size down
// to a floor; the renderer then simply runs more, shorter iterations.
static int wavefront_ensure(ne_b200_ctx* ctx, int lane, uint32_t nSlots) {
	ne_host_span span_("  wavefront_ensure");
	if (ctx->wf[lane] && ctx->wf[lane]->nSlots >= nSlots) return NE_B200_OK;
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (ctx->laneStream[lane]) NE_CUDA_OK(cudaStreamSynchronize(ctx->laneStream[lane]));
	wavefront_free_state(ctx->wf[lane]);
	ctx->wf[lane] = nullptr;
	const uint32_t floorSlots = std::min(nSlots, 1u << 16);
	for (uint32_t n = nSlots;; n = std::max(floorSlots, n / 2)) {
		ne_wavefront_state* w = nullptr;
		cudaError_t e = wavefront_build(ctx, n, &w);
		if (e == cudaSuccess) {
			ctx->wf[lane] = w;
			return NE_B200_OK;
		}
		cudaGetLastError();  // clear the sticky-less allocation error
		if (e != cudaErrorMemoryAllocation || n == floorSlots) {
			set_error(std::string("wavefront pool allocation (") + std::to_string(n) + " slots): " + cudaGetErrorString(e));
			return e == cudaErrorMemoryAllocation ? NE_B200_ERR_NOMEM : NE_B200_ERR_CUDA;
		}
	}
}

static uint32_t env_u32(const char* name, uint32_t dflt) {
	const char* e = getenv(name);
	return e ? (uint32_t)strtoul(e, nullptr, 10) : dflt;
}

// The launches of ONE wavefront iteration on stream `st`, followed by the plan of the next one. Stages the scene cannot
// need are not launched at all (no medium: no tracking / scatter / transmittance kernels; no shadeable surface: no
// k_wf_surface). `mark(kind)` closes a stage in the host-driven loop (a CUDA event); inside the graph the kernels stamp
// the device clock themselves (stage_stamp) and `mark` does nothing.
static uint32_t launches_per_iteration(const WfVariant& V) {
	// generate, commit, extend, plan + (track, scatter, tr with media) + (surface with shadeable surfaces) + (shadow, and trfind
	// with media, unless k_wf_scatter<FUSE> answers its queries in place and no other kernel queues any)
	const bool cast = V.trace || !V.fuse || V.surfaces;
	return 4u + (V.media ? 3u : 0u) + (V.surfaces ? 1u : 0u) + (cast ? (V.media ? 2u : 1u) : 0u);
}

template <class MARK>
static void launch_iteration(cudaStream_t st, const ne_wavefront_state* w, const WfBuf& b, WfParams P, const WfVariant& V, DCounters* counters,
                             cudaGraphConditionalHandle loop, int inGraph, volatile uint32_t* hostDone, MARK&& mark) {
	const int G = w->gridBlocks, B = 256;
	const int GR = w->smCount * NE_TRACE_BLOCKS;
	const int GT = w->smCount * NE_TRACK_BLOCKS;  // persistent tracking kernels: exactly the resident blocks
	const int GS = w->smCount;                    // ... or one 1024-thread block per SM with the majorant tables in shared memory
	const size_t smem = size_t(V.sm);
	int prev = STAGE_OTHER;  // the plan that precedes every iteration
	auto next = [&](int kind) {
		P.prevStage = V.stamps ? prev : -1;
		prev = kind;
	};
	// camera rays are coherent: the grid-stride kernel is as fast as a trace job (measured)
	next(STAGE_OTHER);
	if (V.genBlocks >= 4) k_wf_generate<4><<<G, B, 0, st>>>(b, P);
	else if (V.genBlocks == 3) k_wf_generate<3><<<G, B, 0, st>>>(b, P);
	else k_wf_generate<2><<<G, B, 0, st>>>(b, P);
	next(STAGE_OTHER);
	k_wf_commit<<<1, 1, 0, st>>>(b, counters, P.prevStage, P.genToVol);
	mark(STAGE_OTHER);
	next(STAGE_TRACE);
	if (V.trace) k_wf_trace<ExtendJob><<<GR, 256, 0, st>>>(b, P);
	else k_wf_extend<<<G, B, 0, st>>>(b, P);
	mark(STAGE_TRACE);
	if (V.media) {
		next(STAGE_VOLUME);
		if (!V.brick) k_wf_track<TRACK_GLOBAL, NE_TRACK_THREADS, NE_TRACK_BLOCKS><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
		else if (V.sm && V.skip) k_wf_track<TRACK_SKIP_SM, 1024, 1><<<GS, 1024, smem, st>>>(b, P);
		else if (V.sm) k_wf_track<TRACK_BRICK_SM, 1024, 1><<<GS, 1024, smem, st>>>(b, P);
		else if (V.skip) k_wf_track<TRACK_SKIP, NE_TRACK_THREADS, NE_TRACK_BLOCKS><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
		else k_wf_track<TRACK_BRICK, NE_TRACK_THREADS, NE_TRACK_BLOCKS><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
		mark(STAGE_VOLUME);
		next(STAGE_SHADE);
#define NE_SCATTER(FUSE, FAST)                                                                                     \
	{                                                                                                              \
		if (V.lightSet == PRIM_POINT) k_wf_scatter<FUSE, PRIM_POINT, FAST><<<G, B, 0, st>>>(b, P);                 \
		else if (V.lightSet == PRIM_RECTANGLE) k_wf_scatter<FUSE, PRIM_RECTANGLE, FAST><<<G, B, 0, st>>>(b, P);    \
		else k_wf_scatter<FUSE, -1, FAST><<<G, B, 0, st>>>(b, P);                                                  \
	}
		if (V.fuse && V.fastShade) NE_SCATTER(true, true)
		else if (V.fuse) NE_SCATTER(true, false)
		else if (V.fastShade) NE_SCATTER(false, true)
		else NE_SCATTER(false, false)
#undef NE_SCATTER
	}
	if (V.surfaces) {
		next(STAGE_SHADE);
		if (V.lightSet == PRIM_POINT) k_wf_surface<PRIM_POINT><<<G, B, 0, st>>>(b, P);
		else if (V.lightSet == PRIM_RECTANGLE) k_wf_surface<PRIM_RECTANGLE><<<G, B, 0, st>>>(b, P);
		else if (V.lightSet == PRIM_SPHERE) k_wf_surface<PRIM_SPHERE><<<G, B, 0, st>>>(b, P);
		else k_wf_surface<-1><<<G, B, 0, st>>>(b, P);
	}
	mark(STAGE_SHADE);
	// the shadow / medium-search kernels run only if some kernel queued such requests: k_wf_scatter<FUSE> (mesh-free scenes)
	// answers its own in place, k_wf_surface queues them
	const bool cast = V.trace || !V.fuse || V.surfaces;
	if (cast) {
		next(STAGE_TRACE);
		if (V.trace) k_wf_trace<ShadowJob><<<GR, 256, 0, st>>>(b, P);
		else k_wf_shadow<<<G, B, 0, st>>>(b, P);
	}
	if (V.media) {
		if (cast) {
			next(STAGE_TRACE);
			if (V.trace) k_wf_trace<TrFindJob><<<GR, 256, 0, st>>>(b, P);
			else k_wf_trfind<<<G, B, 0, st>>>(b, P);
		}
		mark(STAGE_TRACE);
		next(STAGE_VOLUME);
		if (!V.brick) k_wf_tr<TRACK_GLOBAL, NE_TRACK_THREADS, NE_TRACK_BLOCKS><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
		else if (V.sm && V.skip) k_wf_tr<TRACK_SKIP_SM, 1024, 1><<<GS, 1024, smem, st>>>(b, P);
		else if (V.sm) k_wf_tr<TRACK_BRICK_SM, 1024, 1><<<GS, 1024, smem, st>>>(b, P);
		else if (V.skip) k_wf_tr<TRACK_SKIP, NE_TRACK_THREADS, NE_TRACK_BLOCKS><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
		else k_wf_tr<TRACK_BRICK, NE_TRACK_THREADS, NE_TRACK_BLOCKS><<<GT, NE_TRACK_THREADS, 0, st>>>(b, P);
		mark(STAGE_VOLUME);
	} else mark(STAGE_TRACE);
	next(STAGE_OTHER);
	k_wf_plan<<<1, 1, 0, st>>>(b, loop, inGraph, hostDone, launches_per_iteration(V), P.prevStage);
}

// Optional: keep the largest brick pool resident in L2 (persisting access-policy window on the launching stream; kernel
// nodes captured from the stream inherit it) so that streaming path records cannot evict voxel data.
static void set_l2_window(ne_b200_ctx* ctx, cudaStream_t st, bool on) {
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof(v));
	if (on && ctx->l2Pool) {
		cudaDeviceProp prop;
		if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return;
		size_t bytes = std::min<size_t>(ctx->l2PoolBytes, size_t(prop.accessPolicyMaxWindowSize));
		size_t carve = std::min<size_t>(bytes, size_t(prop.persistingL2CacheMaxSize));
		cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
		v.accessPolicyWindow.base_ptr = const_cast<void*>(ctx->l2Pool);
		v.accessPolicyWindow.num_bytes = bytes;
		v.accessPolicyWindow.hitRatio = bytes ? float(std::min(1.0, double(carve) / double(bytes))) : 0.0f;
		v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	}
	cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
	cudaGetLastError();
}

// first plan -> WHILE(loop handle){ iteration; plan } -> finish, instantiated once per variant. The loop condition is set on the
// device by k_wf_plan, so a whole render is ONE graph launch: no host round trip, no empty iteration after the last one,
// and ne_b200_render returns while the GPU works.
static int graph_build(ne_b200_ctx* ctx, ne_wavefront_state* w, const WfParams& P, const WfVariant& V) {
	ne_host_span span_("  graph_build");
	graph_free(w);
	NE_CUDA_OK(cudaGraphCreate(&w->graph, 0));
	cudaGraphConditionalHandle loop;
	NE_CUDA_OK(cudaGraphConditionalHandleCreate(&loop, w->graph, 1, cudaGraphCondAssignDefault));
	WfBuf b = w->b;
	int inGraph = 1;
	volatile uint32_t* noHost = nullptr;
	uint32_t perIter = launches_per_iteration(V);
	int noStage = -1;
	void* planArgs[] = {&b, &loop, &inGraph, &noHost, &perIter, &noStage};
	cudaKernelNodeParams kp;
	memset(&kp, 0, sizeof(kp));
	kp.func = reinterpret_cast<void*>(k_wf_plan);
	kp.gridDim = dim3(1);
	kp.blockDim = dim3(1);
	kp.kernelParams = planArgs;
	cudaGraphNode_t nPlan, nLoop, nFinish;
	NE_CUDA_OK(cudaGraphAddKernelNode(&nPlan, w->graph, nullptr, 0, &kp));
	cudaGraphNodeParams np = {};
	np.type = cudaGraphNodeTypeConditional;
	np.conditional.handle = loop;
	np.conditional.type = cudaGraphCondTypeWhile;
	np.conditional.size = 1;
	NE_CUDA_OK(cudaGraphAddNode(&nLoop, w->graph, &nPlan, 1, &np));
	cudaGraph_t body = np.conditional.phGraph_out[0];
	set_l2_window(ctx, w->capStream, V.l2persist != 0);
	NE_CUDA_OK(cudaStreamBeginCaptureToGraph(w->capStream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
	launch_iteration(w->capStream, w, b, P, V, ctx->dCounters, loop, 1, nullptr, [](int) {});
	cudaGraph_t captured = nullptr;
	cudaError_t ce = cudaStreamEndCapture(w->capStream, &captured);
	if (ce != cudaSuccess) {
		set_error(std::string("render graph capture: ") + cudaGetErrorString(ce));
		cudaGetLastError();
		return NE_B200_ERR_CUDA;
	}
	DCounters* counters = ctx->dCounters;
	int foldTimes = V.foldTimes;
	void* finArgs[] = {&b, &counters, &foldTimes};
	kp.func = reinterpret_cast<void*>(k_wf_finish);
	kp.kernelParams = finArgs;
	NE_CUDA_OK(cudaGraphAddKernelNode(&nFinish, w->graph, &nLoop, 1, &kp));
	NE_CUDA_OK(cudaGraphInstantiate(&w->exec, w->graph, 0));
	w->built = V;
	w->haveGraph = true;
	return NE_B200_OK;
}

// Camera-ray culling: the pixel rectangle outside of which NO camera ray can hit anything. Every hittable instance's
// world-space bounds (ctx->boundCorners, built at upload) are projected through the lens onto the film: a ray through film
// point F(u,v) and lens point o = position + offset reaches P = o + s (o - F), s > 0 (Camera::getRayPassingThrough,
// core/Camera.cpp:140-144, d = -normalize(F - o)), so P is seen at F = o - (P - o) / s with s fixed by F lying in the film
// plane. The projection of a convex set is the hull of its projected corners, for every lens offset inside the square
// that holds the lens disk; the rectangle is their bounding box plus a pixel of margin. A path through a pixel outside it
// misses every instance: it carries no radiance (the scene has no light that shines on rays that miss, else cullable is
// false), needs no slot and no ray - it is only counted. Whole frame when any corner is not safely in front of the lens.
static void cull_rect(const ne_b200_ctx* ctx, int* rx0, int* ry0, int* rw, int* rh) {
	const int W = ctx->W, H = ctx->H;
	*rx0 = 0; *ry0 = 0; *rw = W; *rh = H;
	if (!ctx->cullable || getenv("NE_B200_NO_CULL")) return;
	const DCamera& c = ctx->cam;
	auto dot3 = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
	const double Hv[3] = {c.horizontal.x, c.horizontal.y, c.horizontal.z}, Vv[3] = {c.vertical.x, c.vertical.y, c.vertical.z};
	const double n[3] = {Hv[1] * Vv[2] - Hv[2] * Vv[1], Hv[2] * Vv[0] - Hv[0] * Vv[2], Hv[0] * Vv[1] - Hv[1] * Vv[0]};
	const double hh = dot3(Hv, Hv), vv = dot3(Vv, Vv);
	if (!(hh > 0) || !(vv > 0)) return;
	double u0 = INFINITY, u1 = -INFINITY, v0 = INFINITY, v1 = -INFINITY;
	const double R = std::fabs(double(c.lens_radius)) * 1.001;
	for (int k = 0; k < 4; k++) {  // the corners of the square around the lens disk
		const double sx = (k & 1) ? R : -R, sy = (k & 2) ? R : -R;
		const double o[3] = {c.position.x + sx * c.side.x + sy * c.up.x, c.position.y + sx * c.side.y + sy * c.up.y, c.position.z + sx * c.side.z + sy * c.up.z};
		const double oll[3] = {o[0] - c.lower_left.x, o[1] - c.lower_left.y, o[2] - c.lower_left.z};
		const double den = dot3(n, oll);
		if (!(std::fabs(den) > 0)) return;
		for (size_t i = 0; i + 2 < ctx->boundCorners.size(); i += 3) {
			const double Po[3] = {ctx->boundCorners[i] - o[0], ctx->boundCorners[i + 1] - o[1], ctx->boundCorners[i + 2] - o[2]};
			const double s = dot3(n, Po) / den;
			// the corner must be well in front of the lens (s = distance along the view axis in units of the film distance)
			if (!(s > 1e-3) || !std::isfinite(s)) return;
			const double F[3] = {oll[0] - Po[0] / s, oll[1] - Po[1] / s, oll[2] - Po[2] / s};  // F - lower_left
			const double u = dot3(F, Hv) / hh, v = dot3(F, Vv) / vv;
			if (!std::isfinite(u) || !std::isfinite(v)) return;
			u0 = std::min(u0, u); u1 = std::max(u1, u); v0 = std::min(v0, v); v1 = std::max(v1, v);
		}
	}
	if (!(u0 <= u1)) {  // nothing hittable at all: one pixel keeps the bookkeeping uniform
		*rw = 1; *rh = 1;
		return;
	}
	// pixel x covers u in [x / W, (x + 1) / W): one pixel of margin plus a relative epsilon for the fp32 arithmetic of the kernel
	const double eps = 1e-4;
	int x0 = int(std::floor((u0 - eps) * W)) - 1, x1 = int(std::ceil((u1 + eps) * W)) + 1;
	int y0 = int(std::floor((v0 - eps) * H)) - 1, y1 = int(std::ceil((v1 + eps) * H)) + 1;
	x0 = std::max(0, std::min(W, x0)); x1 = std::max(x0, std::min(W, x1));
	y0 = std::max(0, std::min(H, y0)); y1 = std::max(y0, std::min(H, y1));
	if (W % 8 == 0 && H % 4 == 0) {  // keep the 8x4 tiling of the camera rays
		x0 &= ~7; y0 &= ~3;
		x1 = std::min(W, (x1 + 7) & ~7); y1 = std::min(H, (y1 + 3) & ~3);
	}
	if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = 1; y1 = 1; }  // off-screen: as above
	*rx0 = x0; *ry0 = y0; *rw = x1 - x0; *rh = y1 - y0;
}

// One lane's render of samples [sppBegin, sppEnd) on stream `st` (see wavefront_render). prepareOnly: everything up to the
// launch - the lane's pool and its render graph, rebuilt if the scene or a knob changed - so that a two-lane render can get
// BOTH graphs ready before it launches either: instantiating a graph waits for the work already running on the device (the
// second lane of the first frame after an upload used to start when the first was over: 1.3 ms per end-to-end frame).
static int wavefront_render_lane(ne_b200_ctx* ctx, int lane, int nLanes, cudaStream_t st, int sppBegin, int sppEnd, int bounces, uint64_t seed, uint32_t flags,
                                 bool hostLoopAsked, bool prepareOnly = false) {
	int rx0, ry0, rw, rh;
	cull_rect(ctx, &rx0, &ry0, &rw, &rh);
	const unsigned long long work = (unsigned long long)rw * rh * (unsigned long long)(sppEnd - sppBegin);
	if (!prepareOnly) ctx->pathsCulled += ((unsigned long long)ctx->W * ctx->H - (unsigned long long)rw * rh) * (unsigned long long)(sppEnd - sppBegin);
	uint32_t pool = std::max(1024u, env_u32("NE_B200_POOL", 1u << 26) / uint32_t(nLanes));
	uint32_t nSlots = uint32_t(std::min<unsigned long long>(work, pool));
	int rc = wavefront_ensure(ctx, lane, nSlots);
	if (rc) return rc;
	ne_wavefront_state* w = ctx->wf[lane];
	nSlots = std::min(nSlots, w->nSlots);  // a pool kept from a larger render is used up to what this one asks for
	w->b.nSlots = nSlots;
	WfParams P;
	memset(&P, 0, sizeof(P));
	P.s = ctx->scene;
	P.W = ctx->W;
	P.H = ctx->H;
	P.budget = int(std::max(1u, env_u32("NE_B200_TRACK_BUDGET", 128)));
	P.moves = int(std::max(1u, env_u32("NE_B200_TRACK_MOVES", 4)));  // brick crossings per lane between two candidate phases
	P.refill = int(std::min(32u, std::max(1u, env_u32("NE_B200_TRACK_REFILL", 20))));  // refill a warp once this many lanes are idle
	P.walkBudget = int(std::max(1u, env_u32("NE_B200_WALK_BUDGET", 24)));  // inner-node visits per walking lane between two votes
	P.walkRefill = int(std::min(32u, std::max(1u, env_u32("NE_B200_WALK_REFILL", 12))));  // refill a trace warp once this many lanes are not walking
	// a walk over its event budget is cut and queued for the next pass. Cutting by the event count alone keeps the set of paths
	// a pure function of (seed, pixel, sample) - which walks get cut does not depend on timing. NE_B200_TRACK_CUT_ALWAYS=0 cuts
	// only once the kernel's queue has run dry (the tail): the same speed on C2 (26.8 vs 26.7 ms), but run-to-run different paths
	P.cutAlways = env_u32("NE_B200_TRACK_CUT_ALWAYS", 1) ? 1 : 0;
	P.counters = ctx->dCounters;
	WfVariant V;
	memset(&V, 0, sizeof(V));
	V.sceneGen = ctx->sceneGen;
	V.W = ctx->W;
	V.H = ctx->H;
	V.nSlots = nSlots;
	// persistent trace kernels when there are BVHs to walk (NE_B200_TRACE=0/1 overrides)
	V.trace = env_u32("NE_B200_TRACE", ctx->nMeshes > 0 ? 1 : 0) != 0;
	// without meshes, k_wf_scatter traces its own continuation ray (NE_B200_FUSE=0/1 overrides)
	V.fuse = ctx->nMeshes == 0 && env_u32("NE_B200_FUSE", 1) != 0;
	V.brick = !(flags & NE_B200_RENDER_GLOBAL_MAJORANT);
	V.genBlocks = int(env_u32("NE_B200_GEN_BLOCKS", V.trace ? 4 : 3));  // resident blocks k_wf_generate is compiled for (C2: 2: 33.9, 3: 33.2, 4: 33.6 ms)
	// empty-space skipping in the tracking walks pays where a good part of a brick table is far from any density
	// (ctx->skipWorthwhile, decided at upload; NE_B200_SKIP=0/1 overrides)
	V.skip = env_u32("NE_B200_SKIP", ctx->skipWorthwhile ? 1 : 0) != 0;
	// majorant tables in shared memory when they fit beside the staged scene tables (and the scene is one stage_scene
	// stages: NE_CACHE_* entries); V.sm holds the dynamic shared-memory bytes. NE_B200_SMEM_MAJ=0 keeps them in global memory.
	{
		const size_t room = w->smemOptin > sizeof(SceneCache) + 1024 ? w->smemOptin - sizeof(SceneCache) - 1024 : 0;
		const bool staged = P.s.n_inst <= NE_CACHE_INST && P.s.n_mat <= NE_CACHE_MAT && P.s.n_vol <= NE_CACHE_VOL;
		const bool fits = V.brick && staged && ctx->majTableBytes > 0 && ctx->majTableBytes <= room;
		V.sm = (fits && env_u32("NE_B200_SMEM_MAJ", 1) != 0) ? int(ctx->majTableBytes) : 0;
	}
	// shading kernels specialised for the scene's light set (NE_B200_LIGHT_SET=-1 forces the generic ones)
	V.lightSet = getenv("NE_B200_LIGHT_SET") ? atoi(getenv("NE_B200_LIGHT_SET")) : ctx->lightSet;
	V.fastShade = env_u32("NE_B200_EXACT_SHADING", 0) ? 0 : 1;  // FAST medium shading in k_wf_scatter (ne_device.cuh)
	V.stamps = getenv("NE_B200_NO_STAGE_TIMES") == nullptr;
	V.media = ctx->scene.has_medium ? 1 : 0;
	V.surfaces = ctx->nSurfaces > 0 ? 1 : 0;
	V.budget = P.budget; V.refill = P.refill; V.moves = P.moves; V.walkBudget = P.walkBudget; V.walkRefill = P.walkRefill; V.cutAlways = P.cutAlways;
	V.l2persist = env_u32("NE_B200_L2_PERSIST", 0) ? 1 : 0;
	P.genToVol = V.genToVol = (ctx->onlyGridMedia && env_u32("NE_B200_GEN_TO_VOL", 1)) ? 1 : 0;
	V.foldTimes = lane == 0 ? 1 : 0;  // concurrent lanes: lane 0's stage clock stands for the render
	if (V.sm) {
		static const void* smKernels[] = {(const void*)k_wf_track<TRACK_BRICK_SM, 1024, 1>, (const void*)k_wf_track<TRACK_SKIP_SM, 1024, 1>,
		                                  (const void*)k_wf_tr<TRACK_BRICK_SM, 1024, 1>, (const void*)k_wf_tr<TRACK_SKIP_SM, 1024, 1>};
		for (const void* k : smKernels) NE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, V.sm));
	}

	WfDyn dyn;
	dyn.cam = ctx->cam;
	dyn.accum = ctx->accum;
	dyn.seed = seed;
	dyn.sppBegin = sppBegin;
	dyn.bounces = bounces;
	dyn.rx0 = rx0; dyn.ry0 = ry0; dyn.rw = rw; dyn.rh = rh;

	// ---- production: one graph launch, no host involvement until ne_b200_wait
	const bool hostLoop = hostLoopAsked || w->graphBroken;
	if (!hostLoop) {
		if (!w->haveGraph || memcmp(&w->built, &V, sizeof(V)) != 0) {
			rc = graph_build(ctx, w, P, V);
			if (rc) {
				// a driver that cannot build the conditional graph still renders: host-driven loop over the same kernels
				// (NE_B200_REQUIRE_GRAPH=1 makes it an error instead: tests)
				fprintf(stderr, "narvalengine_b200: render graph unavailable (%s); using the host-driven loop\n", ne_b200_last_error());
				graph_free(w);
				cudaGetLastError();
				if (env_u32("NE_B200_REQUIRE_GRAPH", 0)) return rc;
				w->graphBroken = true;
			}
		}
		if (prepareOnly) return NE_B200_OK;
		if (w->haveGraph) {
			k_wf_init<<<1, 1, 0, st>>>(w->b, work, dyn);
			NE_CUDA_OK(cudaGetLastError());
			NE_CUDA_OK(cudaGraphLaunch(w->exec, st));
			ctx->renderPending = true;
			return NE_B200_OK;
		}
	}

	if (prepareOnly) return NE_B200_OK;
	// ---- host-driven loop over the same kernels (NE_B200_HOST_LOOP=1: per-stage CUDA events)
	set_l2_window(ctx, st, V.l2persist != 0);
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
	const bool timeStages = V.stamps != 0;
	cudaGraphConditionalHandle noLoop = 0;
	k_wf_init<<<1, 1, 0, st>>>(w->b, work, dyn);
	*w->hostDone = 0;
	k_wf_plan<<<1, 1, 0, st>>>(w->b, noLoop, 0, w->devDone, launches_per_iteration(V), -1);
	bool done = false;
	WfVariant VH = V;
	VH.stamps = 0;  // this mode's stage times are the CUDA events below
	while (!done) {
		// a few iterations per host poll; finished iterations cost only empty launches
		for (int k = 0; k < 4; k++) {
			cudaEvent_t last = timeStages ? ev() : nullptr;
			launch_iteration(st, w, w->b, P, VH, ctx->dCounters, noLoop, 0, w->devDone, [&](int kind) {
				if (!timeStages) return;
				cudaEvent_t e = ev();
				spans.push_back({last, e, kind});
				last = e;
			});
		}
		NE_CUDA_OK(cudaStreamSynchronize(st));
		NE_CUDA_OK(cudaGetLastError());
		done = *w->hostDone != 0;
		if (timeStages) {
			for (const Span& s : spans) {
				float ms = 0;
				cudaEventElapsedTime(&ms, s.a, s.b);
				(s.kind == STAGE_TRACE ? ctx->msExtend : s.kind == STAGE_VOLUME ? ctx->msVolume : s.kind == STAGE_SHADE ? ctx->msShade : ctx->msOther) += ms;
				ctx->msRender += ms;
			}
			spans.clear();
			evUsed = 0;
		}
	}
	// iterations, launches and the overflow flag go through the same device counters as in the graph (the stage times of this
	// mode are the CUDA events above: the stamps were not launched, so k_wf_finish adds only the idle tail to "other")
	k_wf_finish<<<1, 1, 0, st>>>(w->b, ctx->dCounters, 0);
	ctx->renderPending = true;
	return NE_B200_OK;
}

// Samples [sppBegin, sppEnd) of every pixel. The batch is split into TWO independent halves (different sample indices: Philox
// keys make them independent paths) that run as two render graphs on two streams at once: while one lane's kernel drains
// its tail - a persistent tracking kernel waiting for its longest walks, a nearly empty late iteration - the other lane's
// blocks take the idle SMs. That matters most where a GPU's share of a frame is small (8 GPUs x 8 spp of a 1080p frame: 7
// iterations x 9 kernels in 4 ms). Both lanes splat into the same accumulation buffer (atomic adds) and count into the same
// counters. NE_B200_LANES=1 keeps one lane; the host-driven loop and single-sample batches always do.
int wavefront_render(ne_b200_ctx* ctx, int sppBegin, int sppEnd, int bounces, uint64_t seed, uint32_t flags) {
	if ((unsigned long long)ctx->W * ctx->H * (unsigned long long)(sppEnd - sppBegin) == 0 || bounces == 0) return NE_B200_OK;
	const bool hostLoop = env_u32("NE_B200_HOST_LOOP", 0) != 0;
	// the render's device time (ms_render): an event pair around it on the context's stream, folded in by ne_b200_get_counters.
	// (The host-driven loop adds up its own per-stage events instead.)
	cudaEvent_t spanEnd = nullptr;
	if (!hostLoop) {
		// at most 64 pairs wait for a reader: a caller that never fetches the counters re-uses the last pair (and loses those times)
		const size_t k = std::min<size_t>(ctx->spansPending, 63);
		while (ctx->spanEvents.size() < 2 * (k + 1)) {
			cudaEvent_t e;
			NE_CUDA_OK(cudaEventCreate(&e));
			ctx->spanEvents.push_back(e);
		}
		NE_CUDA_OK(cudaEventRecord(ctx->spanEvents[2 * k], ctx->stream));
		spanEnd = ctx->spanEvents[2 * k + 1];
	}
	struct SpanClose {  // records the closing event on every way out
		ne_b200_ctx* c; cudaEvent_t e;
		~SpanClose() { if (e) { cudaEventRecord(e, c->stream); c->spansPending = std::min<size_t>(c->spansPending + 1, 64); } }
	} spanClose{ctx, spanEnd};
	int lanes = int(std::min(2u, std::max(1u, env_u32("NE_B200_LANES", 2))));
	{
		// two lanes pay when each has enough paths to fill the GPU on its own; for small batches the doubled number of
		// (small) launches costs more than the filled tails give back (measured on the C2 frame: 64 spp 27.0 -> 26.7 ms, 32 spp
		// 14.80 -> 14.66, 16 spp 7.84 -> 8.00, 8 spp 4.34 -> 4.69 ms; and with two devices driven from one process, 32 spp per
		// device ran 14.6 -> 17-21 ms with two lanes each): one lane below 24 Mi paths
		int rx0, ry0, rw, rh;
		cull_rect(ctx, &rx0, &ry0, &rw, &rh);
		const unsigned long long work = (unsigned long long)rw * rh * (unsigned long long)(sppEnd - sppBegin);
		if (work < (unsigned long long)env_u32("NE_B200_LANES_MIN_WORK", 24u << 20)) lanes = 1;
	}
	if (hostLoop || sppEnd - sppBegin < 2) lanes = 1;
	if (lanes == 1) return wavefront_render_lane(ctx, 0, 1, ctx->stream, sppBegin, sppEnd, bounces, seed, flags, hostLoop);
	if (!ctx->laneFork) {
		NE_CUDA_OK(cudaEventCreateWithFlags(&ctx->laneFork, cudaEventDisableTiming));
		for (int l = 0; l < 2; l++) {
			NE_CUDA_OK(cudaStreamCreateWithFlags(&ctx->laneStream[l], cudaStreamNonBlocking));
			NE_CUDA_OK(cudaEventCreateWithFlags(&ctx->laneJoin[l], cudaEventDisableTiming));
		}
	}
	const int mid = sppBegin + (sppEnd - sppBegin + 1) / 2;
	for (int l = 0; l < 2; l++) {
		int rc = wavefront_render_lane(ctx, l, 2, ctx->laneStream[l], l == 0 ? sppBegin : mid, l == 0 ? mid : sppEnd, bounces, seed, flags, false, true);
		if (rc) return rc;
	}
	NE_CUDA_OK(cudaEventRecord(ctx->laneFork, ctx->stream));  // after whatever the caller queued before (clear, uploads)
	for (int l = 0; l < 2; l++) {
		NE_CUDA_OK(cudaStreamWaitEvent(ctx->laneStream[l], ctx->laneFork, 0));
		int rc = wavefront_render_lane(ctx, l, 2, ctx->laneStream[l], l == 0 ? sppBegin : mid, l == 0 ? mid : sppEnd, bounces, seed, flags, false);
		if (rc) return rc;
		NE_CUDA_OK(cudaEventRecord(ctx->laneJoin[l], ctx->laneStream[l]));
		NE_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->laneJoin[l], 0));
	}
	return NE_B200_OK;
}

}  // namespace ne
