// ne_wavefront.cu — the production renderer: OfflineEngine::renderTile's pixel x sample loops
// (core/OfflineEngine.cpp:61-71) over the whole frame as a WAVEFRONT of SoA path records.
//
//   pool      N path slots (SoA float4 arrays, 56 B of state each) that are refilled with new camera samples as
//             paths terminate, so the wavefront stays full until the work runs out
//   queues    arrays of slot indices: extend -> {volume, surface} -> next extend; free slots; all pushes are
//             warp-aggregated (one atomicAdd per warp per queue)
//   kernels   plan (1 thread: queue bookkeeping) · generate (camera rays) · extend (Scene::intersectScene fold, BVH)
//             · volume (delta tracking + phase + next-event setup) · surface (GGX shading + next-event setup)
//             · shadow (visibilityTr requests) · tr (intersectTr + ratio tracking requests)
//   output    fp32 atomicAdd splats into the context's linear accumulation buffer
//
// Every kernel is a grid-stride loop over a device-side count with a fixed grid of (SM count x resident blocks), so
// no host round trip is needed to size launches; the host only polls a mapped "done" word.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ne_ctx.h"
#include "ne_integrator.cuh"

using namespace ne;

namespace {

struct WfCounts {
	uint32_t extend, next, vol, surf, freeN, shadow, tr, gen;
	unsigned long long workNext, workTotal;
	uint32_t done, pad;
};

struct WfBuf {
	// path record: A=(o.xyz,d.x) B=(d.yz,T.xy) C=(T.z,pixel,sample,dim) D=(bounce|guard<<8, nee)
	float4 *pA, *pB, *pC;
	uint2* pD;
	// hit record: A=(p.xyz,tNear) B=(n.xyz,tFar) C=(u,v,inst,prim)
	float4 *hA, *hB, *hC;
	// shadow request: A=(o.xyz,C.x) B=(C.yz,w.xy) C=(w.z,pixel)
	float4 *sA, *sB;
	float2* sC;
	// transmittance request: A=(o.xyz,d.x) B=(d.yz,w.xy) C=(w.z,pixel,sample,stream)
	float4 *tA, *tB, *tC;
	uint32_t *qExtend, *qNext, *qVol, *qSurf, *qFree;
	WfCounts* c;
};

struct WfParams {
	DScene s;
	DCamera cam;
	float* accum;
	int W, H, sppBegin, bounces;
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

__device__ __forceinline__ void flush_stats_wf(const Stats& st, DCounters* c, unsigned paths) {
	unsigned m = __activemask();
	unsigned lane = threadIdx.x & 31;
	unsigned leader = __ffs(m) - 1;
#define NE_FLUSH(field, val)                                                   \
	{                                                                          \
		unsigned v = __reduce_add_sync(m, (unsigned)(val));                    \
		if (lane == leader && v) atomicAdd(&c->field, (unsigned long long)v);  \
	}
	NE_FLUSH(paths, paths)
	NE_FLUSH(extend_rays, st.extend_rays)
	NE_FLUSH(shadow_rays, st.shadow_rays)
	NE_FLUSH(delta_steps, st.delta_steps)
	NE_FLUSH(ratio_steps, st.ratio_steps)
	NE_FLUSH(brick_visits, st.brick_visits)
	NE_FLUSH(bvh_nodes, st.bvh_nodes)
	NE_FLUSH(tri_tests, st.tri_tests)
	NE_FLUSH(prim_tests, st.prim_tests)
	NE_FLUSH(scatter_events, st.scatter_events)
	NE_FLUSH(surface_events, st.surface_events)
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
};
__device__ __forceinline__ PathRec load_path(const WfBuf& b, uint32_t slot) {
	float4 A = b.pA[slot], B = b.pB[slot], C = b.pC[slot];
	uint2 D = b.pD[slot];
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
	return r;
}
__device__ __forceinline__ void store_path(const WfBuf& b, uint32_t slot, const PathRec& r) {
	b.pA[slot] = make_float4(r.ps.ray.o.x, r.ps.ray.o.y, r.ps.ray.o.z, r.ps.ray.d.x);
	b.pB[slot] = make_float4(r.ps.ray.d.y, r.ps.ray.d.z, r.ps.T.x, r.ps.T.y);
	b.pC[slot] = make_float4(r.ps.T.z, __uint_as_float(r.pixel), __uint_as_float(r.sample), __uint_as_float(r.dim));
	b.pD[slot] = make_uint2(uint32_t(r.ps.bounce) | (uint32_t(r.ps.guard) << 8), r.ps.nee);
}
__device__ __forceinline__ Hit load_hit(const WfBuf& b, uint32_t slot) {
	float4 A = b.hA[slot], B = b.hB[slot], C = b.hC[slot];
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
	b.hA[slot] = make_float4(h.p.x, h.p.y, h.p.z, h.tNear);
	b.hB[slot] = make_float4(h.n.x, h.n.y, h.n.z, h.tFar);
	b.hC[slot] = make_float4(h.u, h.v, __int_as_float(h.inst), __int_as_float(h.prim));
}

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

// One thread: retire the finished iteration (next -> extend, clear stage queues) and plan the refill.
// `flip` tells which of the two extend buffers is current; the host alternates it.
__global__ void k_wf_plan(WfBuf b, volatile uint32_t* hostDone) {
	WfCounts& c = *b.c;
	c.extend = c.next;
	c.next = 0;
	c.vol = c.surf = c.shadow = c.tr = 0;
	unsigned long long remaining = c.workTotal - c.workNext;
	uint32_t gen = uint32_t(remaining < c.freeN ? remaining : c.freeN);
	c.gen = gen;
	c.freeN -= gen;
	c.done = (c.extend == 0 && gen == 0) ? 1u : 0u;
	if (hostDone) *hostDone = c.done;
}

// OfflineEngine.cpp:64-67 — sample jitter + Camera::getRayPassingThrough for `gen` new paths into free slots.
__global__ void __launch_bounds__(256) k_wf_generate(WfBuf b, WfParams P) {
	const uint32_t gen = b.c->gen;
	const uint32_t freeN = b.c->freeN;
	const uint32_t extendBase = b.c->extend;
	const unsigned long long workBase = b.c->workNext;
	const uint32_t npix = uint32_t(P.W) * uint32_t(P.H);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < gen; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qFree[freeN + gen - 1 - i];
		unsigned long long w = workBase + i;
		uint32_t pixel = uint32_t(w % npix);
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
		store_path(b, slot, r);
		b.qExtend[extendBase + i] = slot;
	}
}
// One thread: publish the refill (after generate has read the old counts).
__global__ void k_wf_commit(WfBuf b, DCounters* counters) {
	WfCounts& c = *b.c;
	c.extend += c.gen;
	c.workNext += c.gen;
	atomicAdd(&counters->paths, (unsigned long long)c.gen);
	c.gen = 0;
}

// Scene::intersectScene for every path of the extend queue + classify (Li :187-193, :244-260).
__global__ void __launch_bounds__(256) k_wf_extend(WfBuf b, WfParams P) {
	const uint32_t n = b.c->extend;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = b.qExtend[i];
		PathRec r = load_path(b, slot);
		Hit h;
		st.extend_rays++;
		bool did = intersect_scene(P.s, r.ps.ray, h, float(NE_EPSILON12), INFINITY, st);
		QueueSink sink;
		sink.accum = P.accum;
		sink.pixel = r.pixel;
		int kind = classify_hit(P.s, did, h, r.ps, sink);
		if (kind == HIT_TERMINATE) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			store_hit(b, slot, h);
			if (kind == HIT_VOLUME) b.qVol[warp_push(&b.c->vol)] = slot;
			else b.qSurf[warp_push(&b.c->surf)] = slot;
		}
	}
	flush_stats_wf(st, P.counters, 0);
}

template <bool BRICKMAJ, bool VOLUME>
__global__ void __launch_bounds__(256) k_wf_shade(WfBuf b, WfParams P) {
	const uint32_t n = VOLUME ? b.c->vol : b.c->surf;
	const uint32_t* q = VOLUME ? b.qVol : b.qSurf;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t slot = q[i];
		PathRec r = load_path(b, slot);
		Hit h = load_hit(b, slot);
		PhiloxRng rng;
		rng.init(P.seed, r.pixel, r.sample, r.dim);
		QueueSink sink;
		sink.b = &b;
		sink.accum = P.accum;
		sink.pixel = r.pixel;
		sink.sample = r.sample;
		int next = VOLUME ? shade_volume<PhiloxRng, BRICKMAJ>(P.s, r.ps, h, rng, sink, st) : shade_surface<PhiloxRng>(P.s, r.ps, h, rng, sink, st);
		if (next == PATH_NEXT_BOUNCE) r.ps.bounce++;
		if (next == PATH_DONE || r.ps.bounce >= P.bounces) {
			b.qFree[warp_push(&b.c->freeN)] = slot;
		} else {
			r.dim = rng.dim;
			store_path(b, slot, r);
			b.qNext[warp_push(&b.c->next)] = slot;
		}
	}
	flush_stats_wf(st, P.counters, 0);
}

// visibilityTr requests: splat the weight when nothing or an emitter is hit first.
__global__ void __launch_bounds__(256) k_wf_shadow(WfBuf b, WfParams P) {
	const uint32_t n = b.c->shadow;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 A = b.sA[i], B = b.sB[i];
		float2 C = b.sC[i];
		PhiloxRng dummy;
		float vis = visibility_tr<PhiloxRng, false, true>(P.s, V3(A.x, A.y, A.z), V3(A.w, B.x, B.y), dummy, st);
		if (vis != 0) splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * vis);
	}
	flush_stats_wf(st, P.counters, 0);
}

// intersectTr requests: walk through surfaces to the first medium, ratio-track through it, splat weight * Tr.
template <bool BRICKMAJ>
__global__ void __launch_bounds__(256) k_wf_tr(WfBuf b, WfParams P) {
	const uint32_t n = b.c->tr;
	Stats st;
	st.clear();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 A = b.tA[i], B = b.tB[i], C = b.tC[i];
		Ray ray;
		ray.o = V3(A.x, A.y, A.z);
		ray.d = V3(A.w, B.x, B.y);
		PhiloxRng rng;
		rng.init(P.seed, __float_as_uint(C.y), __float_as_uint(C.z), 0u, __float_as_uint(C.w));
		float Tr;
		bool found = intersect_tr<PhiloxRng, false, BRICKMAJ>(P.s, ray, Tr, rng, st);
		if (found && Tr != 0) splat(P.accum, __float_as_uint(C.y), V3(B.z, B.w, C.x) * V3(Tr));
	}
	flush_stats_wf(st, P.counters, 0);
}

}  // namespace

struct ne_wavefront_state {
	WfBuf b{};
	uint32_t nSlots = 0;
	std::vector<void*> allocs;
	uint32_t* hostDone = nullptr;  // pinned, mapped
	uint32_t* devDone = nullptr;
	int gridBlocks = 0;
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
	A(pA) A(pB) A(pC) A(pD) A(hA) A(hB) A(hC) A(sA) A(sB) A(sC) A(tA) A(tB) A(tC) A(qExtend) A(qNext) A(qVol) A(qSurf) A(qFree)
#undef A
	if ((rc = wf_alloc(w, &b.c, 1))) return rc;
	NE_CUDA_OK(cudaHostAlloc(&w->hostDone, sizeof(uint32_t), cudaHostAllocMapped));
	NE_CUDA_OK(cudaHostGetDevicePointer(&w->devDone, w->hostDone, 0));
	cudaDeviceProp prop;
	NE_CUDA_OK(cudaGetDeviceProperties(&prop, ctx->device));
	w->gridBlocks = prop.multiProcessorCount * 8;  // 148 SMs x 8 resident 256-thread blocks
	return NE_B200_OK;
}

int wavefront_render(ne_b200_ctx* ctx, int sppBegin, int sppEnd, int bounces, uint64_t seed, uint32_t flags) {
	const unsigned long long work = (unsigned long long)ctx->W * ctx->H * (unsigned long long)(sppEnd - sppBegin);
	if (work == 0 || bounces == 0) return NE_B200_OK;
	uint32_t pool = 1u << 21;
	if (const char* e = getenv("NE_B200_POOL")) pool = std::max(1024u, (uint32_t)strtoul(e, nullptr, 10));
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
	P.seed = seed;
	P.counters = ctx->dCounters;
	const bool brick = !(flags & NE_B200_RENDER_GLOBAL_MAJORANT);
	cudaStream_t st = ctx->stream;
	const int G = w->gridBlocks, B = 256;

	// event pool for per-stage device times (volume / extend / shade), resolved after the loop
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
			if (iter & 1) std::swap(b.qExtend, b.qNext);
			k_wf_plan<<<1, 1, 0, st>>>(b, w->devDone);
			k_wf_generate<<<G, B, 0, st>>>(b, P);
			k_wf_commit<<<1, 1, 0, st>>>(b, ctx->dCounters);
			cudaEvent_t e0 = timeStages ? ev() : nullptr;
			k_wf_extend<<<G, B, 0, st>>>(b, P);
			cudaEvent_t e1 = timeStages ? ev() : nullptr;
			if (brick) k_wf_shade<true, true><<<G, B, 0, st>>>(b, P);
			else k_wf_shade<false, true><<<G, B, 0, st>>>(b, P);
			cudaEvent_t e2 = timeStages ? ev() : nullptr;
			k_wf_shade<true, false><<<G, B, 0, st>>>(b, P);
			cudaEvent_t e3 = timeStages ? ev() : nullptr;
			k_wf_shadow<<<G, B, 0, st>>>(b, P);
			cudaEvent_t e4 = timeStages ? ev() : nullptr;
			if (brick) k_wf_tr<true><<<G, B, 0, st>>>(b, P);
			else k_wf_tr<false><<<G, B, 0, st>>>(b, P);
			cudaEvent_t e5 = timeStages ? ev() : nullptr;
			if (timeStages) {
				spans.push_back({e0, e1, 0});
				spans.push_back({e1, e2, 1});
				spans.push_back({e2, e3, 2});
				spans.push_back({e3, e4, 0});
				spans.push_back({e4, e5, 1});
			}
			ctx->kernelLaunches += 8;
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
