// ne_api.cu — the extern "C" boundary of include/ne_b200.h: context, scene flattening + upload to HBM, render entry
// points, framebuffer read-back, counters, and the single-function test hooks. No CPU fallback: every compute entry
// point launches CUDA kernels and returns NE_B200_ERR_CUDA when that is impossible.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>

#include "ne_ctx.h"
#include "ne_integrator.cuh"
#include "ne_host.h"

using namespace ne;

namespace ne {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
}  // namespace ne

namespace {

// The stream of the context a test hook runs on (set by check_ctx): hook uploads are ordered on it, like the hook's kernel.
thread_local cudaStream_t g_hookStream = nullptr;

template <class T>
struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	~DevBuf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t count) {
		n = count;
		return cudaMalloc(&p, std::max<size_t>(1, count) * sizeof(T));
	}
	// on the context's (non-blocking) stream: the legacy default stream the blocking cudaMemcpy uses does not order with it
	cudaError_t upload(const T* h, size_t count) {
		cudaError_t e = alloc(count);
		if (e != cudaSuccess || !count) return e;
		return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, g_hookStream);
	}
	cudaError_t download(T* h) const {
		if (!n) return cudaSuccess;
		cudaError_t e = cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, g_hookStream);
		return e != cudaSuccess ? e : cudaStreamSynchronize(g_hookStream);
	}
};

__device__ __forceinline__ void flush_stats(const Stats& st, DCounters* c, unsigned paths) {
	unsigned m = __activemask();
	unsigned lane = threadIdx.x & 31;
	unsigned leader = __ffs(m) - 1;
#define NE_FLUSH(field, val)                                        \
	{                                                               \
		unsigned v = __reduce_add_sync(m, (unsigned)(val));         \
		if (lane == leader && v) atomicAdd(&c->field, (unsigned long long)v); \
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

// ---------------------------------------------------------------------------------------------------------------
// Kernels: megakernel render (one thread per pixel, loops over its samples), resolve, test hooks
// ---------------------------------------------------------------------------------------------------------------
template <bool BRICKMAJ, bool FAST>  // FAST: medium shading as k_wf_scatter<.., FASTSH> does it (ne_device.cuh "FAST medium shading")
__global__ void __launch_bounds__(128) k_render_mega(DScene s, DCamera cam, float* accum, int W, int H, int sppBegin, int sppEnd, int bounces,
                                                     uint64_t seed, DCounters* counters) {
	// 8x4 pixel tiles per warp for ray coherence
	int tilesX = (W + 7) / 8;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int tx = warp % tilesX, ty = warp / tilesX;
	int x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);
	Stats st;
	st.clear();
	unsigned paths = 0;
	if (x < W && y < H) {
		uint32_t pixel = uint32_t(W) * y + x;
		V3 sum(0.0f);
		for (int smp = sppBegin; smp < sppEnd; smp++) {
			PhiloxRng rng;
			rng.init(seed, pixel, uint32_t(smp));
			float u = float(float(x) + rng.next()) / float(W);  // OfflineEngine.cpp:65-66
			float v = float(float(y) + rng.next()) / float(H);
			Ray r = camera_ray(cam, u, v, rng);
			sum = sum + li_path<PhiloxRng, false, BRICKMAJ, FAST>(s, r, bounces, rng, st);
			paths++;
		}
		accum[3 * size_t(pixel)] += sum.x;
		accum[3 * size_t(pixel) + 1] += sum.y;
		accum[3 * size_t(pixel) + 2] += sum.z;
	}
	flush_stats(st, counters, paths);
}

__global__ void k_resolve(const float* accum, float invSamples, size_t n, float* linear, float* tonemapped) {
	size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float c = accum[i] * invSamples;
	if (linear) linear[i] = c;
	if (tonemapped) tonemapped[i] = tonemap1(c);
}

// Adaptive stopping: A holds the sums of the even-numbered sample batches (nA samples per pixel), B of the odd ones (nB).
// Two independent estimates of the same frame: ((A/nA - B/nB) / 2)^2 estimates the squared error of their average, so
// the mean over pixels and channels of that over (mean^2 + eps) estimates the rel-MSE (SURVEY 8d's metric, eps = 1e-4) of
// the frame rendered so far against the converged one. out[0] += the block's partial sum (double).
__global__ void __launch_bounds__(256) k_adaptive_error(const float* A, const float* B, float invA, float invB, float invN, size_t n, double* out) {
	double acc = 0;
	for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
		float a = A[i], b = B[i];
		float d = 0.5f * (a * invA - b * invB);
		float m = (a + b) * invN;
		acc += double(d * d / (m * m + 1e-4f));
	}
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
	__shared__ double part[8];
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double t = 0;
		for (int k = 0; k < 8; k++) t += part[k];
		atomicAdd(out, t);
	}
}
__global__ void k_add_into(float* A, const float* B, size_t n) {
	size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n) A[i] += B[i];
}

__global__ void k_test_intersect(DScene s, int n, const float* o, const float* d, float tMin, float tMax, ne_b200_hit* out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Ray r;
	r.o = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
	r.d = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
	Stats st;
	st.clear();
	Hit h;
	bool did = intersect_scene(s, r, h, tMin, tMax, st);
	ne_b200_hit q;
	memset(&q, 0, sizeof(q));
	q.hit = did ? 1 : 0;
	q.instance = -1;
	if (did) {
		q.hit_point[0] = h.p.x; q.hit_point[1] = h.p.y; q.hit_point[2] = h.p.z;
		q.normal[0] = h.n.x; q.normal[1] = h.n.y; q.normal[2] = h.n.z;
		q.uv[0] = h.u; q.uv[1] = h.v;
		q.t_near = h.tNear; q.t_far = h.tFar;
		q.instance = h.inst;
		int mi = s.inst[h.inst].material;
		q.is_light = (mi >= 0 && s.mat[mi].has_light) ? 1 : 0;
		q.primitive = h.prim;
	}
	out[i] = q;
}

__global__ void k_test_camera(DCamera cam, int n, const float* xy, const float* tape, float* o, float* d) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	TapeRng rng;
	rng.init(tape + 2 * i, 2);
	Ray r = camera_ray(cam, xy[2 * i], xy[2 * i + 1], rng);
	o[3 * i] = r.o.x; o[3 * i + 1] = r.o.y; o[3 * i + 2] = r.o.z;
	d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z;
}

__device__ Hit make_hit(const ne_b200_hit& q) {
	Hit h;
	h.p = V3(q.hit_point[0], q.hit_point[1], q.hit_point[2]);
	h.n = V3(q.normal[0], q.normal[1], q.normal[2]);
	h.u = q.uv[0]; h.v = q.uv[1];
	h.tNear = q.t_near; h.tFar = q.t_far;
	h.inst = q.instance;
	h.prim = q.primitive;
	return h;
}

template <bool FAST>  // FAST: the production versions of the medium's phase function (instance must carry a medium)
__global__ void k_test_bsdf(DScene s, int n, int instance, const float* in, const float* sc, const float* nrm, const float* uv,
                            const float* tape, float* eval, float* pdf, float* sampled) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const DMaterial& m = s.mat[s.inst[instance].material];
	Hit h;
	h.p = V3(0.0f);
	h.n = V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
	h.u = uv ? uv[2 * i] : 0.0f;
	h.v = uv ? uv[2 * i + 1] : 0.0f;
	h.tNear = h.tFar = 0;
	h.inst = instance;
	h.prim = 0;
	V3 a(in[3 * i], in[3 * i + 1], in[3 * i + 2]), b(sc[3 * i], sc[3 * i + 1], sc[3 * i + 2]);
	if (eval) {
		V3 e = FAST ? bsdf_eval<1, FAST>(s, m, a, b, h) : bsdf_eval(s, m, a, b, h);
		eval[3 * i] = e.x; eval[3 * i + 1] = e.y; eval[3 * i + 2] = e.z;
	}
	if (pdf) pdf[i] = FAST ? bsdf_pdf<1, FAST>(s, m, a, b, h.n, h) : bsdf_pdf(s, m, a, b, h.n, h);
	if (tape && sampled) {
		TapeRng rng;
		rng.init(tape + 2 * i, 2);
		V3 w = FAST ? bsdf_sample<1, FAST>(s, m, a, h.n, h, rng) : bsdf_sample(s, m, a, h.n, h, rng);
		sampled[3 * i] = w.x; sampled[3 * i + 1] = w.y; sampled[3 * i + 2] = w.z;
	}
}

__global__ void k_test_grid_tr(DScene s, int n, int instance, const float* o, const float* d, const float* tn, const float* tf, const float* tape,
                               int stride, float* tr, int* used) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const DInstance& in = s.inst[instance];
	const DMaterial& m = s.mat[in.material];
	Ray r;
	r.o = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
	r.d = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
	TapeRng rng;
	rng.init(tape + size_t(stride) * i, stride);
	Stats st;
	st.clear();
	tr[i] = grid_tr<TapeRng, false>(in, m, s.vol[m.volume], r, tn[i], tf[i], rng, st);
	if (used) used[i] = rng.overflow ? -1 : rng.pos;
}

__global__ void k_test_grid_sample(DScene s, int n, int instance, const float* o, const float* d, const float* tn, const float* tf,
                                   const float* tape, int stride, float* T, float* so, float* sd, int* used) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const DInstance& in = s.inst[instance];
	const DMaterial& m = s.mat[in.material];
	Ray r;
	r.o = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
	r.d = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
	TapeRng rng;
	rng.init(tape + size_t(stride) * i, stride);
	Stats st;
	st.clear();
	Hit h;
	h.p = r.o; h.n = V3(0.0f, 1.0f, 0.0f); h.u = h.v = 0; h.tNear = tn[i]; h.tFar = tf[i]; h.inst = instance; h.prim = 0;
	Ray sc;
	V3 a = grid_sample<TapeRng, false>(s, in, m, s.vol[m.volume], r, tn[i], tf[i], h, sc, rng, st);
	T[3 * i] = a.x; T[3 * i + 1] = a.y; T[3 * i + 2] = a.z;
	so[3 * i] = sc.o.x; so[3 * i + 1] = sc.o.y; so[3 * i + 2] = sc.o.z;
	sd[3 * i] = sc.d.x; sd[3 * i + 1] = sc.d.y; sd[3 * i + 2] = sc.d.z;
	if (used) used[i] = rng.overflow ? -1 : rng.pos;
}

template <bool FAST>
__global__ void k_test_li_tape(DScene s, int n, const float* o, const float* d, int bounces, const float* tape, int stride, float* L, int* used) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Ray r;
	r.o = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
	r.d = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
	TapeRng rng;
	rng.init(tape + size_t(stride) * i, stride);
	Stats st;
	st.clear();
	V3 v = li_path<TapeRng, true, false, FAST>(s, r, bounces, rng, st);
	L[3 * i] = v.x; L[3 * i + 1] = v.y; L[3 * i + 2] = v.z;
	if (used) used[i] = rng.overflow ? -1 : rng.pos;
}

template <bool BRICKMAJ>
__global__ void k_test_li_philox(DScene s, int n, const float* o, const float* d, int bounces, uint64_t seed, float* L, DCounters* counters) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	Stats st;
	st.clear();
	if (i < n) {
		Ray r;
		r.o = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
		r.d = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
		PhiloxRng rng;
		rng.init(seed, uint32_t(i), 0u);
		V3 v = li_path<PhiloxRng, false, BRICKMAJ>(s, r, bounces, rng, st);
		L[3 * i] = v.x; L[3 * i + 1] = v.y; L[3 * i + 2] = v.z;
	}
	flush_stats(st, counters, i < n ? 1u : 0u);
}

template <bool FAST>  // FAST: every hit must lie in a medium
__global__ void k_test_one_light(DScene s, int n, const float* dirs, const ne_b200_hit* hits, const float* tape, int stride, float* L, int* used) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Hit h = make_hit(hits[i]);
	Ray in;
	in.d = V3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
	in.o = h.p - in.d;
	TapeRng rng;
	rng.init(tape + size_t(stride) * i, stride);
	Stats st;
	st.clear();
	ImmediateSink<TapeRng, true, false, FAST> sink;
	sink.L = V3(0.0f);
	V3 v = sample_one_light<FAST ? 1 : -1, -1>(s, in, h, rng, sink, 1u, st);
	L[3 * i] = v.x; L[3 * i + 1] = v.y; L[3 * i + 2] = v.z;
	if (used) used[i] = rng.overflow ? -1 : rng.pos;
}

__global__ void k_test_density(DScene s, int n, int instance, const float* pts, float* out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const DVolume& v = s.vol[s.mat[s.inst[instance].material].volume];
	out[i] = interpolated_density(v, ocs_to_gcs(v, V3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])));
}

__global__ void k_test_philox(uint64_t seed, uint32_t pixel, uint32_t sample, int n, float* out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	PhiloxRng rng;
	rng.init(seed, pixel, sample, uint32_t(i));
	out[i] = rng.next();
}

inline int blocks(size_t n, int bs) { return int((n + bs - 1) / bs); }

int check_ctx(ne_b200_ctx* ctx, bool needScene) {
	if (!ctx) { set_error("null context"); return NE_B200_ERR_INVALID; }
	NE_CUDA_OK(cudaSetDevice(ctx->device));
	g_hookStream = ctx->stream;
	if (needScene && !ctx->haveScene) { set_error("no scene uploaded"); return NE_B200_ERR_STATE; }
	return NE_B200_OK;
}

template <class T>
int push_alloc(ne_b200_ctx* ctx, const T* host, size_t count, const T** out) {
	// stream-ordered allocation from the device's default pool (release threshold raised in ne_b200_create): re-uploading
	// a scene recycles the previous scene's blocks instead of going through cudaMalloc / cudaFree (measured: occasional
	// 30-250 ms stalls per upload with the synchronous allocator)
	T* p = nullptr;
	NE_CUDA_OK(cudaMallocAsync(&p, std::max<size_t>(1, count) * sizeof(T), ctx->stream));
	ctx->sceneAllocs.push_back(p);
	if (count) NE_CUDA_OK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));  // `host` may be a temporary
	*out = p;
	return NE_B200_OK;
}

void free_scene(ne_b200_ctx* ctx) {
	for (void* p : ctx->sceneAllocs) cudaFreeAsync(p, ctx->stream);
	ctx->sceneAllocs.clear();
	ctx->haveScene = false;
	memset(&ctx->scene, 0, sizeof(ctx->scene));
}

// Texture::sample on the host (materials/Texture.cpp:37-129), channel `ch` of the nearest texel: the same arithmetic as
// tex_sample / tex_at in ne_device.cuh.
float host_tex_sample(const ne_b200_texture& t, float u, float v, int ch) {
	auto wrap = [](float x, int mode) {
		if (mode == 1) { float f = std::fabs(x - float(int(x))); return f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f); }
		return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
	};
	u = wrap(u, t.wrap_u);
	v = wrap(v, t.wrap_v);
	int x = int(u * t.width), y = int(v * t.height);
	if (x > 0 && x == t.width) x--;
	if (y > 0 && y == t.height) y--;
	size_t index = size_t(t.width) * y + x;
	if (t.format == TEX_RGBA8) return static_cast<const uint8_t*>(t.texels)[4 * index + ch] / 255.0f;
	const float* p = static_cast<const float*>(t.texels);
	int n = t.format == TEX_R32F ? 1 : t.format == TEX_RG32F ? 2 : t.format == TEX_RGB32F ? 3 : 4;
	return ch < n ? p[n * index + ch] : 0.0f;
}

size_t texel_bytes(int format) {
	switch (format) {
	case TEX_R32F: return 4;
	case TEX_RG32F: return 8;
	case TEX_RGB32F: return 12;
	case TEX_RGBA32F: return 16;
	default: return 4;
	}
}

}  // namespace

extern "C" {

const char* ne_b200_last_error(void) { return g_err.c_str(); }

int ne_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int ne_b200_make_transform(const float position[3], const float rotation_deg[3], const float scale[3], float to_world[16], float to_object[16]) {
	if (!position || !rotation_deg || !scale || !to_world || !to_object) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	host_make_transform(position, rotation_deg, scale, to_world, to_object);
	return NE_B200_OK;
}

int ne_b200_camera_make(const float look_from[3], const float look_at[3], const float up[3], float vfov_deg, float aspect, float aperture,
                        float focus_distance, ne_b200_camera* out) {
	if (!look_from || !look_at || !up || !out) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	host_camera_make(look_from, look_at, up, vfov_deg, aspect, aperture, focus_distance, out);
	return NE_B200_OK;
}

int ne_b200_create(int cuda_device, ne_b200_ctx** out) {
	if (!out) { set_error("null out"); return NE_B200_ERR_INVALID; }
	*out = nullptr;
	int n = ne_b200_device_count();
	if (n <= 0) { set_error("no CUDA device visible (this library has no CPU fallback)"); return NE_B200_ERR_CUDA; }
	if (cuda_device < 0 || cuda_device >= n) { set_error("cuda_device out of range"); return NE_B200_ERR_INVALID; }
	NE_CUDA_OK(cudaSetDevice(cuda_device));
	std::unique_ptr<ne_b200_ctx> ctx(new ne_b200_ctx());
	ctx->device = cuda_device;
	NE_CUDA_OK(cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking));
	ctx->stream = ctx->ownStream;
	NE_CUDA_OK(cudaEventCreate(&ctx->evA));
	NE_CUDA_OK(cudaEventCreate(&ctx->evB));
	{
		cudaMemPool_t pool;
		uint64_t keep = UINT64_MAX;  // keep freed scene memory in the pool for the next upload
		if (cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		cudaGetLastError();
	}
	NE_CUDA_OK(cudaMalloc(&ctx->dCounters, sizeof(DCounters)));
	NE_CUDA_OK(cudaMemset(ctx->dCounters, 0, sizeof(DCounters)));
	ctx->spansPending = 0;
	*out = ctx.release();
	return NE_B200_OK;
}

void ne_b200_destroy(ne_b200_ctx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	wavefront_free(ctx);
	free_scene(ctx);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);  // the stream-ordered frees above
	if (ctx->accum) cudaFree(ctx->accum);
	if (ctx->scratch) cudaFree(ctx->scratch);
	if (ctx->pinned) cudaFreeHost(ctx->pinned);
	if (ctx->dCounters) cudaFree(ctx->dCounters);
	if (ctx->evA) cudaEventDestroy(ctx->evA);
	if (ctx->evB) cudaEventDestroy(ctx->evB);
	for (cudaEvent_t e : ctx->spanEvents) cudaEventDestroy(e);
	if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
	delete ctx;
}

int ne_b200_set_stream(ne_b200_ctx* ctx, void* cuda_stream) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	ctx->stream = static_cast<cudaStream_t>(cuda_stream);
	return NE_B200_OK;
}

int ne_b200_scene_upload(ne_b200_ctx* ctx, const ne_b200_scene_desc* d) {
	ne_host_span span_("scene_upload");
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (!d) { set_error("null scene"); return NE_B200_ERR_INVALID; }
	auto t0 = std::chrono::steady_clock::now();
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	free_scene(ctx);

	// ---- validate + materials (SceneReader::processMaterial, io/SceneReader.cpp:67-222)
	std::vector<DMaterial> mats(d->n_materials);
	std::vector<int> envs;  // env_tex of every infiniteAreaLight material
	for (int i = 0; i < d->n_materials; i++) {
		const ne_b200_material& m = d->materials[i];
		DMaterial& o = mats[i];
		memset(&o, 0, sizeof(o));
		o.type = m.type;
		o.albedo_tex = m.albedo_tex; o.roughness_tex = m.roughness_tex; o.metallic_tex = m.metallic_tex; o.normal_tex = m.normal_tex;
		o.has_normal_flag = (m.has_normal_flag && m.normal_tex >= 0) ? 1 : 0;
		for (int k = 0; k < 3; k++) { o.li[k] = m.li[k]; o.sigma_s[k] = m.scattering[k]; o.sigma_a[k] = m.absorption[k]; }
		o.density_mult = m.density_multiplier;
		o.phase = m.phase;
		o.g = m.g;
		o.volume = m.volume;
		o.light_owner = -1;
		auto texOk = [&](int t) { return t < d->n_textures; };
		if (!texOk(m.albedo_tex) || !texOk(m.roughness_tex) || !texOk(m.metallic_tex) || !texOk(m.normal_tex)) {
			set_error("material texture index out of range");
			return NE_B200_ERR_INVALID;
		}
		switch (m.type) {
		case NE_B200_MAT_MICROFACET: o.has_bsdf = 1; break;
		case NE_B200_MAT_EMITTER: o.has_light = 1; break;
		case NE_B200_MAT_VOLUME:
			o.has_bsdf = 1; o.transmissive = 1; o.has_medium = 1;
			// volume < 0: HomogeneousMedia (SceneReader.cpp:208-210, materials/HomogeneousMedia.cpp)
			if (m.volume >= d->n_volumes) { set_error("material volume index out of range"); return NE_B200_ERR_INVALID; }
			break;
		case NE_B200_MAT_DIRECTIONAL:  // SceneReader.cpp:156-168: le = albedo (in `li` here), direction = normalize(-position)
			o.has_light = 1;
			o.directional = 1;
			for (int k = 0; k < 3; k++) o.direction[k] = m.direction[k];
			break;
		case NE_B200_MAT_INFINITE:  // SceneReader.cpp:169-186: InfiniteAreaLight(tex); li and le stay 0
			if (m.env_tex < 0 || m.env_tex >= d->n_textures) { set_error("infiniteAreaLight needs env_tex"); return NE_B200_ERR_INVALID; }
			o.has_light = 1;
			o.infinite = 1;
			o.env = int(envs.size());
			for (int k = 0; k < 3; k++) o.li[k] = 0.0f;
			envs.push_back(m.env_tex);
			break;
		default: set_error("unknown material type"); return NE_B200_ERR_INVALID;
		}
	}

	// ---- fold order (Scene::instancedModels..., Scene::lights...; SceneReader.cpp:224-648, SceneEditor.cpp:2039-2054)
	std::vector<int> models, lights;
	for (int i = 0; i < d->n_primitives; i++) {
		const ne_b200_primitive& p = d->primitives[i];
		if (p.material >= d->n_materials) { set_error("primitive material index out of range"); return NE_B200_ERR_INVALID; }
		if (p.type < 0 || p.type > NE_B200_PRIM_MESH) { set_error("unknown primitive type"); return NE_B200_ERR_INVALID; }
		bool volMat = p.material >= 0 && mats[p.material].has_medium;
		if (p.type == NE_B200_PRIM_VOLUME && !volMat) { set_error("volume primitive needs a volume material"); return NE_B200_ERR_INVALID; }
		if (p.type != NE_B200_PRIM_VOLUME && volMat) { set_error("volume material on a non-volume primitive is not supported"); return NE_B200_ERR_UNSUPPORTED; }
		if (p.type != NE_B200_PRIM_MESH && p.material < 0) { set_error("analytic primitive without material"); return NE_B200_ERR_INVALID; }
		if (p.type == NE_B200_PRIM_MESH && (p.n_triangles < 0 || p.n_vertices < 0 || (p.n_triangles > 0 && (!p.positions || !p.indices)))) {
			set_error("mesh without geometry");
			return NE_B200_ERR_INVALID;
		}
		bool isLight = p.type != NE_B200_PRIM_MESH && p.type != NE_B200_PRIM_VOLUME && mats[p.material].has_light;
		(isLight ? lights : models).push_back(i);
	}
	if (d->sort_and_group)
		std::stable_partition(models.begin(), models.end(), [&](int i) { return !(d->primitives[i].material >= 0 && mats[d->primitives[i].material].has_medium); });
	std::vector<int> fold = models;
	fold.insert(fold.end(), lights.begin(), lights.end());

	// ---- textures
	std::vector<DTexture> texs(d->n_textures);
	for (int i = 0; i < d->n_textures; i++) {
		const ne_b200_texture& t = d->textures[i];
		if (t.width <= 0 || t.height <= 0 || !t.texels || t.format < 0 || t.format > TEX_RGBA8) { set_error("bad texture"); return NE_B200_ERR_INVALID; }
		texs[i].w = t.width; texs[i].h = t.height; texs[i].format = t.format; texs[i].wrap_u = t.wrap_u; texs[i].wrap_v = t.wrap_v;
		const uint8_t* dev;
		rc = push_alloc(ctx, (const uint8_t*)t.texels, size_t(t.width) * t.height * texel_bytes(t.format), &dev);
		if (rc) return rc;
		texs[i].texels = dev;
	}

	// ---- InfiniteAreaLight::InfiniteAreaLight(tex), lights/InfiniteAreaLight.h:20-41 + Distribution2D (Sampling.h:69-94)
	std::vector<DEnvDist> envDists(envs.size());
	for (size_t e = 0; e < envs.size(); e++) {
		const ne_b200_texture& t = d->textures[envs[e]];
		const int w = t.width, h = t.height, nc = h;  // Q27: every conditional is built with n = height
		std::vector<float> img(size_t(w) * h + size_t(nc), 0.0f);  // zero padding where the reference reads past the array
		for (int v = 0; v < h; v++) {
			float vp = float(v) / float(h);
			float sinTheta = float(std::sin(NE_PI * double(float(float(v) + 0.5f)) / double(float(h))));
			for (int u = 0; u < w; u++) {
				float up = float(u) / float(w);
				img[u + size_t(v) * w] = host_tex_sample(t, up, vp, 1);
				img[u + size_t(v) * w] *= sinTheta;
			}
		}
		auto dist1d = [](const float* f, int n, std::vector<float>& cdf) {  // Distribution1D ctor, Sampling.h:12-29
			cdf.assign(n + 1, 0.0f);
			for (int i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + f[i - 1] / n;
			float funcInt = cdf[n];
			if (funcInt == 0) for (int i = 1; i < n + 1; i++) cdf[i] = float(i) / float(n);
			else for (int i = 1; i < n + 1; i++) cdf[i] /= funcInt;
			return funcInt;
		};
		std::vector<float> cFunc(size_t(h) * nc), cCdf(size_t(h) * (nc + 1)), cInt(h), mCdf, tmp;
		for (int v = 0; v < h; v++) {
			memcpy(&cFunc[size_t(v) * nc], &img[size_t(v) * w], nc * sizeof(float));
			cInt[v] = dist1d(&img[size_t(v) * w], nc, tmp);
			memcpy(&cCdf[size_t(v) * (nc + 1)], tmp.data(), (nc + 1) * sizeof(float));
		}
		DEnvDist& o = envDists[e];
		o.tex = envs[e]; o.w = w; o.h = h; o.nc = nc;
		o.mInt = dist1d(cInt.data(), h, mCdf);
		if ((rc = push_alloc(ctx, cFunc.data(), cFunc.size(), &o.cFunc))) return rc;
		if ((rc = push_alloc(ctx, cCdf.data(), cCdf.size(), &o.cCdf))) return rc;
		if ((rc = push_alloc(ctx, cInt.data(), cInt.size(), &o.cInt))) return rc;
		if ((rc = push_alloc(ctx, cInt.data(), cInt.size(), &o.mFunc))) return rc;
		if ((rc = push_alloc(ctx, mCdf.data(), mCdf.size(), &o.mCdf))) return rc;
	}

	// ---- volumes -> brick-sparse grids
	std::vector<DVolume> vols(d->n_volumes);
	for (int i = 0; i < d->n_volumes; i++) {
		const ne_b200_volume& v = d->volumes[i];
		if (v.width <= 0 || v.height <= 0 || v.depth <= 0 || (!v.dense && v.n_leaves > 0 && (!v.leaf_origin || !v.leaf_values))) {
			set_error("bad volume");
			return NE_B200_ERR_INVALID;
		}
		DVolume& o = vols[i];
		if (v.dense) {  // dense grid: copied once, bricked by three kernels (ne_bricks.cu)
			if ((rc = device_build_bricks(ctx, v, o))) return rc;
			continue;
		}
		HostBricks hb;
		host_build_bricks(v, hb);  // OpenVDB leaves: scattered input, bricked on the host
		o.W = hb.W; o.H = hb.H; o.D = hb.D; o.bx = hb.bx; o.by = hb.by; o.bz = hb.bz;
		o.max_density = hb.maxDensity;
		o.inv_max_density = 1.0f / hb.maxDensity;
		// one 8-byte cell per brick: {slot, 1/majorant}; a brick without a record has nothing to collide with
		std::vector<int2> cells(hb.table.size());
		for (size_t b = 0; b < cells.size(); b++) {
			float inv = hb.table[b] >= 0 ? hb.binv[b] : 0.0f;
			cells[b].x = hb.table[b];
			memcpy(&cells[b].y, &inv, 4);
		}
		o.n_slots = int(hb.pool.size() / BRICK_VOX);
		if ((rc = push_alloc(ctx, cells.data(), cells.size(), &o.cells))) return rc;
		if ((rc = push_alloc(ctx, hb.pool.data(), hb.pool.size(), &o.pool))) return rc;
	}
	ctx->skipWorthwhile = false;
	for (DVolume& o : vols)
		if ((rc = device_build_majorants(ctx, o))) return rc;

	// ---- instances + meshes
	std::vector<DInstance> insts(fold.size());
	std::vector<DMesh> meshes;
	std::vector<int> foldOf(d->n_primitives, -1);
	bool hasMedium = false;
	for (size_t f = 0; f < fold.size(); f++) {
		const ne_b200_primitive& p = d->primitives[fold[f]];
		DInstance& o = insts[f];
		memset(&o, 0, sizeof(o));
		memcpy(o.M, p.to_world, 64);
		memcpy(o.Mi, p.to_object, 64);
		o.type = p.type;
		o.material = p.material;
		o.collision = p.collision;
		o.mesh = -1;
		o.radius = p.radius;
		for (int k = 0; k < 3; k++) o.point[k] = p.point[k];
		host_get_scale(p.to_world, o.scale);
		o.desc_index = fold[f];
		foldOf[fold[f]] = int(f);
		if (p.material >= 0 && mats[p.material].has_medium && int(f) < int(models.size())) hasMedium = true;
		if (p.type == NE_B200_PRIM_MESH) {
			for (int t = 0; t < 3 * p.n_triangles; t++)
				if (p.indices[t] >= uint32_t(p.n_vertices)) { set_error("mesh index out of range"); return NE_B200_ERR_INVALID; }
			HostBvh hb;
			host_build_bvh(p.positions, p.n_vertices, p.indices, p.n_triangles, hb);
			DMesh m;
			memset(&m, 0, sizeof(m));
			m.n_tris = p.n_triangles; m.n_verts = p.n_vertices; m.n_nodes = int(hb.nodes.size());
			for (int k = 0; k < 3; k++) { m.bbmin[k] = hb.bbmin[k]; m.bbmax[k] = hb.bbmax[k]; }
			if ((rc = push_alloc(ctx, hb.nodes.data(), hb.nodes.size(), &m.nodes))) return rc;
			const float* triDev;
			if ((rc = push_alloc(ctx, hb.tri.data(), hb.tri.size(), &triDev))) return rc;
			m.tri = reinterpret_cast<const float4*>(triDev);
			if ((rc = push_alloc(ctx, p.positions, size_t(3) * p.n_vertices, &m.pos))) return rc;
			if ((rc = push_alloc(ctx, p.indices, size_t(3) * p.n_triangles, &m.idx))) return rc;
			m.uv = nullptr;
			if (p.uvs && (rc = push_alloc(ctx, p.uvs, size_t(2) * p.n_vertices, &m.uv))) return rc;
			o.mesh = int(meshes.size());
			meshes.push_back(m);
		}
	}
	// Light::primitive = the LAST primitive built with that emitter material (Q7)
	for (int i = 0; i < d->n_primitives; i++) {
		const ne_b200_primitive& p = d->primitives[i];
		if (p.material >= 0 && mats[p.material].has_light && p.type != NE_B200_PRIM_MESH && p.type != NE_B200_PRIM_VOLUME)
			mats[p.material].light_owner = foldOf[i];
	}

	DScene s;
	memset(&s, 0, sizeof(s));
	s.n_inst = int(fold.size());
	s.n_models = int(models.size());
	s.n_lights = int(lights.size());
	s.has_medium = hasMedium ? 1 : 0;
	s.n_mat = int(mats.size());
	s.n_vol = int(vols.size());
	for (size_t f = models.size(); f < fold.size(); f++) {
		if (mats[d->primitives[fold[f]].material].directional) s.n_directional++;
		if (mats[d->primitives[fold[f]].material].infinite) s.n_infinite++;
	}
	if ((rc = push_alloc(ctx, insts.data(), insts.size(), &s.inst))) return rc;
	if ((rc = push_alloc(ctx, mats.data(), mats.size(), &s.mat))) return rc;
	if ((rc = push_alloc(ctx, texs.data(), texs.size(), &s.tex))) return rc;
	if ((rc = push_alloc(ctx, vols.data(), vols.size(), &s.vol))) return rc;
	if ((rc = push_alloc(ctx, meshes.data(), meshes.size(), &s.mesh))) return rc;
	if ((rc = push_alloc(ctx, envDists.data(), envDists.size(), &s.env))) return rc;
	ctx->scene = s;
	ctx->nVolumes = d->n_volumes;
	ctx->nMeshes = int(meshes.size());
	ctx->haveScene = true;
	ctx->sceneGen++;
	// bounds of everything hittable, in world space (camera-ray culling): OCS extent of each primitive kind through M
	ctx->boundCorners.clear();
	ctx->cullable = s.n_directional == 0 && s.n_infinite == 0;
	for (const DInstance& in : insts) {
		if (!in.collision || in.type == PRIM_POINT) continue;  // never hit (InstancedModel.cpp:25, Point.cpp:10-12)
		float lo[3] = {-0.5f, -0.5f, -0.5f}, hi[3] = {0.5f, 0.5f, 0.5f};  // volume proxy cube
		if (in.type == PRIM_RECTANGLE) { lo[2] = hi[2] = 0.0f; }
		else if (in.type == PRIM_SPHERE) { for (int k = 0; k < 3; k++) { lo[k] = -std::fabs(in.radius); hi[k] = std::fabs(in.radius); } }
		else if (in.type == PRIM_MESH) {
			if (in.mesh < 0) { ctx->cullable = false; continue; }
			for (int k = 0; k < 3; k++) { lo[k] = meshes[in.mesh].bbmin[k]; hi[k] = meshes[in.mesh].bbmax[k]; }
		}
		for (int c = 0; c < 8; c++) {
			const float p[3] = {(c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]};
			for (int r = 0; r < 3; r++) {
				float w = in.M[r] * p[0] + in.M[4 + r] * p[1] + in.M[8 + r] * p[2] + in.M[12 + r];
				if (!std::isfinite(w)) ctx->cullable = false;
				ctx->boundCorners.push_back(w);
			}
		}
	}
	{
		// the scene's light set: every light instance's Light::primitive (Q7: the material's light_owner) of one analytic kind,
		// all of them DiffuseLights, no medium without a grid (the specialised scatter kernel drops that branch too)
		int set = -2;
		for (size_t f = models.size(); f < fold.size(); f++) {
			const DMaterial& lm = mats[insts[f].material];
			const int kind = (lm.directional || lm.infinite || lm.light_owner < 0) ? -1 : insts[lm.light_owner].type;
			const int k2 = (kind == PRIM_RECTANGLE || kind == PRIM_SPHERE || kind == PRIM_POINT) ? kind : -1;
			set = set == -2 ? k2 : (set == k2 ? set : -1);
		}
		for (const DMaterial& m : mats)
			if (m.has_medium && m.volume < 0) set = -1;
		ctx->lightSet = set < 0 ? -1 : set;
	}
	ctx->nSurfaces = 0;
	for (const DInstance& in : insts)
		if (in.material >= 0 && mats[in.material].has_bsdf && !mats[in.material].transmissive) ctx->nSurfaces++;
	// every surviving camera path enters the volume queue: nothing shadeable but grid media (k_wf_generate then needs one reservation
	// per survivor, not two)
	ctx->onlyGridMedia = ctx->nSurfaces == 0 && !vols.empty();
	for (const DMaterial& m : mats)
		if (m.has_medium && m.volume < 0) ctx->onlyGridMedia = false;
	ctx->majTableBytes = 0;
	ctx->l2Pool = nullptr;
	ctx->l2PoolBytes = 0;
	for (const DVolume& o : vols) {
		ctx->majTableBytes += (size_t(o.bx + 2) * (o.by + 2) * (o.bz + 2) * sizeof(unsigned short) + 15) & ~size_t(15);
		const size_t pb = size_t(o.n_slots) * BRICK_VOX * sizeof(float);
		if (pb > ctx->l2PoolBytes) { ctx->l2PoolBytes = pb; ctx->l2Pool = o.pool; }
	}
	ctx->msUpload = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return NE_B200_OK;
}

int ne_b200_camera_set(ne_b200_ctx* ctx, const ne_b200_camera* c) {
	if (!ctx || !c) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	auto v = [](const float* p) { return V3(p[0], p[1], p[2]); };
	ctx->cam.position = v(c->position);
	ctx->cam.lower_left = v(c->lower_left);
	ctx->cam.horizontal = v(c->horizontal);
	ctx->cam.vertical = v(c->vertical);
	ctx->cam.side = v(c->side);
	ctx->cam.up = v(c->up);
	ctx->cam.lens_radius = c->lens_radius;
	ctx->haveCamera = true;
	return NE_B200_OK;
}

int ne_b200_clear(ne_b200_ctx* ctx) {
	ne_host_span span_("clear");
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (ctx->accum) NE_CUDA_OK(cudaMemsetAsync(ctx->accum, 0, size_t(ctx->W) * ctx->H * 3 * sizeof(float), ctx->stream));
	ctx->samples = 0;
	return NE_B200_OK;
}

int ne_b200_render(ne_b200_ctx* ctx, int width, int height, int spp_begin, int spp_end, int bounces, uint64_t seed, uint32_t flags) {
	ne_host_span span_("render (call)");
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if (!ctx->haveCamera) { set_error("no camera set"); return NE_B200_ERR_STATE; }
	if (width <= 0 || height <= 0 || spp_end < spp_begin || spp_begin < 0 || bounces < 0) { set_error("bad render arguments"); return NE_B200_ERR_INVALID; }
	if (bounces > 65535) { set_error("more than 65535 bounces (the path record keeps the bounce count in 16 bits)"); return NE_B200_ERR_INVALID; }
	if (!ctx->accum || width != ctx->W || height != ctx->H) {
		NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		if (ctx->accum) cudaFree(ctx->accum);
		ctx->accum = nullptr;
		NE_CUDA_OK(cudaMalloc(&ctx->accum, size_t(width) * height * 3 * sizeof(float)));
		ctx->W = width;
		ctx->H = height;
		NE_CUDA_OK(cudaMemsetAsync(ctx->accum, 0, size_t(width) * height * 3 * sizeof(float), ctx->stream));
		ctx->samples = 0;
	}
	if (spp_end == spp_begin) return NE_B200_OK;
	if (flags & NE_B200_RENDER_MEGAKERNEL) {
		NE_CUDA_OK(cudaEventRecord(ctx->evA, ctx->stream));
		int tilesX = (width + 7) / 8, tilesY = (height + 3) / 4;
		size_t threads = size_t(tilesX) * tilesY * 32;
		int bs = 128;
		const char* exactShading = getenv("NE_B200_EXACT_SHADING");  // as in the wavefront renderer: the check renderer takes the same decisions
		const bool fast = !(exactShading && atoi(exactShading) != 0);
#define NE_MEGA(BM, FAST) \
	k_render_mega<BM, FAST><<<blocks(threads, bs), bs, 0, ctx->stream>>>(ctx->scene, ctx->cam, ctx->accum, width, height, spp_begin, spp_end, bounces, seed, ctx->dCounters)
		if (flags & NE_B200_RENDER_GLOBAL_MAJORANT) {
			if (fast) NE_MEGA(false, true); else NE_MEGA(false, false);
		} else {
			if (fast) NE_MEGA(true, true); else NE_MEGA(true, false);
		}
#undef NE_MEGA
		ctx->kernelLaunches++;
		NE_CUDA_OK(cudaGetLastError());
		NE_CUDA_OK(cudaEventRecord(ctx->evB, ctx->stream));
		NE_CUDA_OK(cudaEventSynchronize(ctx->evB));  // the debug path keeps its CUDA-event time
		float ms = 0;
		NE_CUDA_OK(cudaEventElapsedTime(&ms, ctx->evA, ctx->evB));
		ctx->msRender += ms;
	} else {
		// asynchronous: the whole render is enqueued as one CUDA graph launch (ne_wavefront.cu); ne_b200_wait joins
		rc = wavefront_render(ctx, spp_begin, spp_end, bounces, seed, flags);
		if (rc) return rc;
	}
	ctx->samples += spp_end - spp_begin;
	return NE_B200_OK;
}

int ne_b200_wait(ne_b200_ctx* ctx) {
	ne_host_span span_("wait");
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	NE_CUDA_OK(cudaGetLastError());
	if (ctx->renderPending) {
		ctx->renderPending = false;
		unsigned long long overflow = 0;
		NE_CUDA_OK(cudaMemcpy(&overflow, &ctx->dCounters->overflow, sizeof(overflow), cudaMemcpyDeviceToHost));
		if (overflow) {
			cudaMemset(&ctx->dCounters->overflow, 0, sizeof(overflow));
			set_error("wavefront request array overflow (internal capacity invariant violated); the frame is incomplete");
			return NE_B200_ERR_STATE;
		}
	}
	return NE_B200_OK;
}

int ne_b200_accum_buffer(ne_b200_ctx* ctx, void** device_ptr, size_t* n_floats, int* samples_accumulated) {
	if (!ctx) { set_error("null context"); return NE_B200_ERR_INVALID; }
	if (!ctx->accum) { set_error("nothing rendered yet"); return NE_B200_ERR_STATE; }
	if (device_ptr) *device_ptr = ctx->accum;
	if (n_floats) *n_floats = size_t(ctx->W) * ctx->H * 3;
	if (samples_accumulated) *samples_accumulated = ctx->samples;
	return NE_B200_OK;
}

int ne_b200_set_samples_accumulated(ne_b200_ctx* ctx, int samples) {
	if (!ctx || samples < 0) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	ctx->samples = samples;
	return NE_B200_OK;
}

int ne_b200_accum_download(ne_b200_ctx* ctx, float* sums, int* samples_accumulated) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (!ctx->accum) { set_error("nothing rendered yet"); return NE_B200_ERR_STATE; }
	if (!sums) { set_error("null buffer"); return NE_B200_ERR_INVALID; }
	NE_CUDA_OK(cudaMemcpyAsync(sums, ctx->accum, size_t(ctx->W) * ctx->H * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (samples_accumulated) *samples_accumulated = ctx->samples;
	return NE_B200_OK;
}

int ne_b200_accum_upload(ne_b200_ctx* ctx, int width, int height, const float* sums, int samples_accumulated) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (!sums || width <= 0 || height <= 0 || samples_accumulated < 0) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!ctx->accum || width != ctx->W || height != ctx->H) {
		NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		if (ctx->accum) cudaFree(ctx->accum);
		ctx->accum = nullptr;
		NE_CUDA_OK(cudaMalloc(&ctx->accum, size_t(width) * height * 3 * sizeof(float)));
		ctx->W = width;
		ctx->H = height;
	}
	NE_CUDA_OK(cudaMemcpyAsync(ctx->accum, sums, size_t(width) * height * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	ctx->samples = samples_accumulated;
	return NE_B200_OK;
}

static int read_resolved(ne_b200_ctx* ctx, float* linear, float* tonemapped) {
	ne_host_span span_("read_resolved");
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (!ctx->accum || ctx->samples <= 0) { set_error("nothing rendered yet"); return NE_B200_ERR_STATE; }
	size_t n = size_t(ctx->W) * ctx->H * 3;
	if ((rc = scratch_reserve(ctx, 2 * n * sizeof(float)))) return rc;  // resolved frames live in the context's scratch
	float* lin = static_cast<float*>(ctx->scratch);
	float* tm = lin + n;
	k_resolve<<<blocks(n, 256), 256, 0, ctx->stream>>>(ctx->accum, 1.0f / float(ctx->samples), n, linear ? lin : nullptr, tonemapped ? tm : nullptr);
	ctx->kernelLaunches++;
	NE_CUDA_OK(cudaGetLastError());
	if (linear) NE_CUDA_OK(cudaMemcpyAsync(linear, lin, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	if (tonemapped) NE_CUDA_OK(cudaMemcpyAsync(tonemapped, tm, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NE_B200_OK;
}

int ne_b200_read_linear(ne_b200_ctx* ctx, float* rgb) {
	if (!rgb) { set_error("null buffer"); return NE_B200_ERR_INVALID; }
	return read_resolved(ctx, rgb, nullptr);
}
int ne_b200_read_tonemapped(ne_b200_ctx* ctx, float* rgb) {
	if (!rgb) { set_error("null buffer"); return NE_B200_ERR_INVALID; }
	return read_resolved(ctx, nullptr, rgb);
}

int ne_b200_render_frame(ne_b200_ctx* ctx, const ne_b200_camera* camera, int width, int height, int spp, int bounces, uint64_t seed, uint32_t flags,
                         float* pixels_tonemapped, float* pixels_linear) {
	int rc;
	if (camera && (rc = ne_b200_camera_set(ctx, camera))) return rc;
	if ((rc = ne_b200_render(ctx, width, height, 0, 0, bounces, seed, flags))) return rc;  // (re)allocate
	if ((rc = ne_b200_clear(ctx))) return rc;
	if ((rc = ne_b200_render(ctx, width, height, 0, spp, bounces, seed, flags))) return rc;
	if ((rc = ne_b200_wait(ctx))) return rc;
	if (!pixels_tonemapped && !pixels_linear) return NE_B200_OK;
	return read_resolved(ctx, pixels_linear, pixels_tonemapped);
}

// Progressive rendering with a stopping rule: sample batches go alternately into two accumulation
// buffers; after every pair the GPU estimates the frame's rel-MSE from their difference (k_adaptive_error) and the loop
// stops once it is below the target (and at least spp_min samples are in) or spp_max is reached. The buffers are then
// summed into the context's accumulation buffer, so the frame can still be checkpointed or continued.
int ne_b200_render_adaptive(ne_b200_ctx* ctx, const ne_b200_camera* camera, int width, int height, int spp_min, int spp_max, int spp_batch,
                            float target_rel_mse, int bounces, uint64_t seed, uint32_t flags, float* pixels_tonemapped, float* pixels_linear,
                            ne_b200_adaptive_result* result) {
	int rc;
	if (spp_batch < 1 || spp_min < 0 || spp_max < 2 * spp_batch || spp_min > spp_max || !(target_rel_mse >= 0)) { set_error("bad adaptive arguments (spp_max >= 2 * spp_batch, spp_batch >= 1)"); return NE_B200_ERR_INVALID; }
	if (camera && (rc = ne_b200_camera_set(ctx, camera))) return rc;
	if ((rc = ne_b200_render(ctx, width, height, 0, 0, bounces, seed, flags))) return rc;  // (re)allocate
	if ((rc = ne_b200_clear(ctx))) return rc;
	const size_t n = size_t(width) * height * 3;
	float* A = ctx->accum;
	float* B = nullptr;
	double* dErr = nullptr;
	NE_CUDA_OK(cudaMallocAsync(&B, n * sizeof(float), ctx->stream));
	NE_CUDA_OK(cudaMallocAsync(&dErr, sizeof(double), ctx->stream));
	NE_CUDA_OK(cudaMemsetAsync(B, 0, n * sizeof(float), ctx->stream));
	int done = 0, nA = 0, nB = 0, batches = 0;
	double err = INFINITY;
	bool converged = false;
	rc = NE_B200_OK;
	while (done < spp_max) {
		const int take = std::min(spp_batch, spp_max - done);
		ctx->accum = (batches & 1) ? B : A;  // the renderer splats into whichever buffer the context names
		rc = ne_b200_render(ctx, width, height, done, done + take, bounces, seed, flags);
		ctx->accum = A;
		if (rc) break;
		((batches & 1) ? nB : nA) += take;
		done += take;
		batches++;
		if ((batches & 1) == 0) {  // a complete pair: two estimates with (nearly) the same sample count
			cudaMemsetAsync(dErr, 0, sizeof(double), ctx->stream);
			k_adaptive_error<<<592, 256, 0, ctx->stream>>>(A, B, 1.0f / float(nA), 1.0f / float(nB), 1.0f / float(nA + nB), n, dErr);
			ctx->kernelLaunches++;
			double sum = 0;
			if (cudaMemcpyAsync(&sum, dErr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
			    cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("adaptive error read-back failed"); rc = NE_B200_ERR_CUDA; break; }
			err = sum / double(n);
			if (done >= spp_min && err <= double(target_rel_mse)) { converged = true; break; }
		}
	}
	if (!rc) {
		k_add_into<<<blocks(n, 256), 256, 0, ctx->stream>>>(A, B, n);
		ctx->kernelLaunches++;
	}
	cudaFreeAsync(B, ctx->stream);
	cudaFreeAsync(dErr, ctx->stream);
	ctx->samples = done;
	if (rc) return rc;
	if ((rc = ne_b200_wait(ctx))) return rc;
	if (result) {
		result->spp_rendered = done;
		result->rel_mse_estimate = float(err);
		result->converged = converged ? 1 : 0;
	}
	if (!pixels_tonemapped && !pixels_linear) return NE_B200_OK;
	return read_resolved(ctx, pixels_linear, pixels_tonemapped);
}

int ne_b200_get_counters(ne_b200_ctx* ctx, ne_b200_counters* out) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (!out) { set_error("null out"); return NE_B200_ERR_INVALID; }
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	for (size_t k = 0; k < ctx->spansPending; k++) {  // device time of the asynchronous renders since the last call
		float ms = 0;
		if (cudaEventElapsedTime(&ms, ctx->spanEvents[2 * k], ctx->spanEvents[2 * k + 1]) == cudaSuccess) ctx->msRender += ms;
	}
	ctx->spansPending = 0;
	cudaGetLastError();
	DCounters c;
	NE_CUDA_OK(cudaMemcpy(&c, ctx->dCounters, sizeof(c), cudaMemcpyDeviceToHost));
	memset(out, 0, sizeof(*out));
	out->paths = c.paths + ctx->pathsCulled; out->extend_rays = c.extend_rays; out->shadow_rays = c.shadow_rays; out->delta_steps = c.delta_steps;
	out->ratio_steps = c.ratio_steps; out->brick_visits = c.brick_visits; out->bvh_nodes = c.bvh_nodes; out->tri_tests = c.tri_tests;
	out->prim_tests = c.prim_tests; out->scatter_events = c.scatter_events; out->surface_events = c.surface_events;
	out->wavefront_iterations = c.iterations;
	out->kernel_launches = ctx->kernelLaunches + c.launches;
	// host-side CUDA-event accounts (megakernel, host-driven loop) + the render graph's device-side stage accounts
	const double ns = 1e-6;
	out->ms_extend_kernel = ctx->msExtend + double(c.stage_ns[0]) * ns;
	out->ms_volume_kernel = ctx->msVolume + double(c.stage_ns[1]) * ns;
	out->ms_shade_kernel = ctx->msShade + double(c.stage_ns[2]) * ns;
	out->ms_other_kernel = ctx->msOther + double(c.stage_ns[3]) * ns;
	out->ms_render = ctx->msRender;  // CUDA events around every render (with two overlapped lanes the stage accounts above are lane 0's clock)
	out->ms_upload = ctx->msUpload;
	// SURVEY 8d's algorithmic figures: 8 voxels x 4 B + brick-table entry 4 B + majorant 4 B per tracking step (this
	// layout reads 8 x 4 B + a 4-byte slot + a 2-byte majorant per brick crossing); 36 B of vertex data per triangle test
	out->bytes_per_tracking_step = 40;
	out->bytes_per_bvh_node = sizeof(BvhNode);
	out->bytes_per_triangle = 36;
	out->bytes_per_path_record = uint32_t(ne::wavefront_record_bytes());
	return NE_B200_OK;
}

int ne_b200_counters_reset(ne_b200_ctx* ctx) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	NE_CUDA_OK(cudaMemset(ctx->dCounters, 0, sizeof(DCounters)));
	ctx->spansPending = 0;
	ctx->kernelLaunches = ctx->wavefrontIterations = 0;
	ctx->pathsCulled = 0;
	ctx->msRender = ctx->msVolume = ctx->msExtend = ctx->msShade = ctx->msOther = 0;
	return NE_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Test hooks
// ---------------------------------------------------------------------------------------------------------------
#define NE_UP(buf, ptr, count) NE_CUDA_OK(buf.upload(ptr, count))
#define NE_FINISH()                                  \
	ctx->kernelLaunches++;                           \
	NE_CUDA_OK(cudaGetLastError());                  \
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream))

int ne_b200_test_intersect(ne_b200_ctx* ctx, int n, const float* origins, const float* directions, float t_min, float t_max, ne_b200_hit* out) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if (n < 0 || (n && (!origins || !directions || !out))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> o, d;
	DevBuf<ne_b200_hit> h;
	NE_UP(o, origins, 3 * size_t(n)); NE_UP(d, directions, 3 * size_t(n));
	NE_CUDA_OK(h.alloc(n));
	k_test_intersect<<<blocks(n, 128), 128, 0, ctx->stream>>>(ctx->scene, n, o.p, d.p, t_min, t_max, h.p);
	NE_FINISH();
	NE_CUDA_OK(h.download(out));
	return NE_B200_OK;
}

int ne_b200_test_camera_rays(ne_b200_ctx* ctx, int n, const float* xy, const float* tape, float* origins, float* directions) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (!ctx->haveCamera) { set_error("no camera set"); return NE_B200_ERR_STATE; }
	if (n < 0 || (n && (!xy || !tape || !origins || !directions))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> a, t, o, d;
	NE_UP(a, xy, 2 * size_t(n)); NE_UP(t, tape, 2 * size_t(n));
	NE_CUDA_OK(o.alloc(3 * size_t(n))); NE_CUDA_OK(d.alloc(3 * size_t(n)));
	k_test_camera<<<blocks(n, 128), 128, 0, ctx->stream>>>(ctx->cam, n, a.p, t.p, o.p, d.p);
	NE_FINISH();
	NE_CUDA_OK(o.download(origins)); NE_CUDA_OK(d.download(directions));
	return NE_B200_OK;
}

static int check_instance(ne_b200_ctx* ctx, int instance, bool needVolume, std::vector<DInstance>* hostInst = nullptr) {
	if (instance < 0 || instance >= ctx->scene.n_inst) { set_error("instance out of range"); return NE_B200_ERR_INVALID; }
	DInstance in;
	NE_CUDA_OK(cudaMemcpy(&in, ctx->scene.inst + instance, sizeof(in), cudaMemcpyDeviceToHost));
	if (in.material < 0) { set_error("instance has no material"); return NE_B200_ERR_INVALID; }
	DMaterial m;
	NE_CUDA_OK(cudaMemcpy(&m, ctx->scene.mat + in.material, sizeof(m), cudaMemcpyDeviceToHost));
	if (needVolume && (!m.has_medium || m.volume < 0)) { set_error("instance has no grid medium"); return NE_B200_ERR_INVALID; }
	if (!needVolume && !m.has_bsdf) { set_error("instance has no BSDF"); return NE_B200_ERR_INVALID; }
	return NE_B200_OK;
}

// FAST medium shading in the hooks (ne_b200_test_set_fast_shading): the instance must carry a medium's phase function
static bool instance_is_medium(ne_b200_ctx* ctx, int instance) {
	if (instance < 0 || instance >= ctx->scene.n_inst) return false;
	DInstance in;
	if (cudaMemcpy(&in, ctx->scene.inst + instance, sizeof(in), cudaMemcpyDeviceToHost) != cudaSuccess || in.material < 0) return false;
	DMaterial m;
	if (cudaMemcpy(&m, ctx->scene.mat + in.material, sizeof(m), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
	return m.has_bsdf && m.transmissive;
}

int ne_b200_test_set_fast_shading(ne_b200_ctx* ctx, int on) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	ctx->testFastShading = on != 0;
	return NE_B200_OK;
}

int ne_b200_test_bsdf(ne_b200_ctx* ctx, int n, int instance, const float* incoming, const float* scattered, const float* normals, const float* uvs,
                      const float* tape, float* eval, float* pdf, float* sampled) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if ((rc = check_instance(ctx, instance, false))) return rc;
	if (n < 0 || (n && (!incoming || !scattered || !normals))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> a, b, c, u, t, e, p, s;
	NE_UP(a, incoming, 3 * size_t(n)); NE_UP(b, scattered, 3 * size_t(n)); NE_UP(c, normals, 3 * size_t(n));
	if (uvs) NE_UP(u, uvs, 2 * size_t(n));
	if (tape) NE_UP(t, tape, 2 * size_t(n));
	NE_CUDA_OK(e.alloc(3 * size_t(n))); NE_CUDA_OK(p.alloc(n)); NE_CUDA_OK(s.alloc(3 * size_t(n)));
	if (ctx->testFastShading) {
		if (!instance_is_medium(ctx, instance)) { set_error("fast shading exists for media only"); return NE_B200_ERR_INVALID; }
		k_test_bsdf<true><<<blocks(n, 128), 128, 0, ctx->stream>>>(ctx->scene, n, instance, a.p, b.p, c.p, uvs ? u.p : nullptr, tape ? t.p : nullptr,
		                                                            eval ? e.p : nullptr, pdf ? p.p : nullptr, sampled ? s.p : nullptr);
	} else
		k_test_bsdf<false><<<blocks(n, 128), 128, 0, ctx->stream>>>(ctx->scene, n, instance, a.p, b.p, c.p, uvs ? u.p : nullptr, tape ? t.p : nullptr,
		                                                             eval ? e.p : nullptr, pdf ? p.p : nullptr, sampled ? s.p : nullptr);
	NE_FINISH();
	if (eval) NE_CUDA_OK(e.download(eval));
	if (pdf) NE_CUDA_OK(p.download(pdf));
	if (sampled && tape) NE_CUDA_OK(s.download(sampled));
	return NE_B200_OK;
}

int ne_b200_test_grid_tr(ne_b200_ctx* ctx, int n, int instance, const float* origins, const float* directions, const float* t_near, const float* t_far,
                         const float* tape, int tape_stride, float* tr, int32_t* used) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if ((rc = check_instance(ctx, instance, true))) return rc;
	if (n < 0 || tape_stride <= 0 || (n && (!origins || !directions || !t_near || !t_far || !tape || !tr))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> o, d, a, b, t, r;
	DevBuf<int> u;
	NE_UP(o, origins, 3 * size_t(n)); NE_UP(d, directions, 3 * size_t(n)); NE_UP(a, t_near, n); NE_UP(b, t_far, n);
	NE_UP(t, tape, size_t(n) * tape_stride);
	NE_CUDA_OK(r.alloc(n)); NE_CUDA_OK(u.alloc(n));
	k_test_grid_tr<<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, instance, o.p, d.p, a.p, b.p, t.p, tape_stride, r.p, u.p);
	NE_FINISH();
	NE_CUDA_OK(r.download(tr));
	if (used) NE_CUDA_OK(u.download(used));
	return NE_B200_OK;
}

int ne_b200_test_grid_sample(ne_b200_ctx* ctx, int n, int instance, const float* origins, const float* directions, const float* t_near,
                             const float* t_far, const float* tape, int tape_stride, float* transmittance, float* scattered_o, float* scattered_d,
                             int32_t* used) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if ((rc = check_instance(ctx, instance, true))) return rc;
	if (n < 0 || tape_stride <= 0 || (n && (!origins || !directions || !t_near || !t_far || !tape || !transmittance || !scattered_o || !scattered_d))) {
		set_error("bad argument");
		return NE_B200_ERR_INVALID;
	}
	if (!n) return NE_B200_OK;
	DevBuf<float> o, d, a, b, t, T, so, sd;
	DevBuf<int> u;
	NE_UP(o, origins, 3 * size_t(n)); NE_UP(d, directions, 3 * size_t(n)); NE_UP(a, t_near, n); NE_UP(b, t_far, n);
	NE_UP(t, tape, size_t(n) * tape_stride);
	NE_CUDA_OK(T.alloc(3 * size_t(n))); NE_CUDA_OK(so.alloc(3 * size_t(n))); NE_CUDA_OK(sd.alloc(3 * size_t(n))); NE_CUDA_OK(u.alloc(n));
	k_test_grid_sample<<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, instance, o.p, d.p, a.p, b.p, t.p, tape_stride, T.p, so.p, sd.p, u.p);
	NE_FINISH();
	NE_CUDA_OK(T.download(transmittance)); NE_CUDA_OK(so.download(scattered_o)); NE_CUDA_OK(sd.download(scattered_d));
	if (used) NE_CUDA_OK(u.download(used));
	return NE_B200_OK;
}

int ne_b200_test_li_tape(ne_b200_ctx* ctx, int n, const float* origins, const float* directions, int bounces, const float* tape, int tape_stride,
                         float* radiance, int32_t* used) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if (n < 0 || tape_stride <= 0 || (n && (!origins || !directions || !tape || !radiance))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> o, d, t, L;
	DevBuf<int> u;
	NE_UP(o, origins, 3 * size_t(n)); NE_UP(d, directions, 3 * size_t(n)); NE_UP(t, tape, size_t(n) * tape_stride);
	NE_CUDA_OK(L.alloc(3 * size_t(n))); NE_CUDA_OK(u.alloc(n));
	if (ctx->testFastShading) k_test_li_tape<true><<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, o.p, d.p, bounces, t.p, tape_stride, L.p, u.p);
	else k_test_li_tape<false><<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, o.p, d.p, bounces, t.p, tape_stride, L.p, u.p);
	NE_FINISH();
	NE_CUDA_OK(L.download(radiance));
	if (used) NE_CUDA_OK(u.download(used));
	return NE_B200_OK;
}

int ne_b200_test_li_philox(ne_b200_ctx* ctx, int n, const float* origins, const float* directions, int bounces, uint64_t seed, uint32_t flags,
                           float* radiance) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if (n < 0 || (n && (!origins || !directions || !radiance))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> o, d, L;
	NE_UP(o, origins, 3 * size_t(n)); NE_UP(d, directions, 3 * size_t(n));
	NE_CUDA_OK(L.alloc(3 * size_t(n)));
	if (flags & NE_B200_RENDER_GLOBAL_MAJORANT)
		k_test_li_philox<false><<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, o.p, d.p, bounces, seed, L.p, ctx->dCounters);
	else
		k_test_li_philox<true><<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, o.p, d.p, bounces, seed, L.p, ctx->dCounters);
	NE_FINISH();
	NE_CUDA_OK(L.download(radiance));
	return NE_B200_OK;
}

int ne_b200_test_sample_one_light(ne_b200_ctx* ctx, int n, const float* incoming_dirs, const ne_b200_hit* hits, const float* tape, int tape_stride,
                                  float* radiance, int32_t* used) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if (n < 0 || tape_stride <= 0 || (n && (!incoming_dirs || !hits || !tape || !radiance))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	for (int i = 0; i < n; i++)
		if (hits[i].instance < 0 || hits[i].instance >= ctx->scene.n_inst) { set_error("hit instance out of range"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> d, t, L;
	DevBuf<ne_b200_hit> h;
	DevBuf<int> u;
	NE_UP(d, incoming_dirs, 3 * size_t(n)); NE_UP(h, hits, n); NE_UP(t, tape, size_t(n) * tape_stride);
	NE_CUDA_OK(L.alloc(3 * size_t(n))); NE_CUDA_OK(u.alloc(n));
	if (ctx->testFastShading) {
		int lastChecked = -1;
		for (int i = 0; i < n; i++) {
			if (hits[i].instance == lastChecked) continue;
			if (!instance_is_medium(ctx, hits[i].instance)) { set_error("fast shading exists for media only"); return NE_B200_ERR_INVALID; }
			lastChecked = hits[i].instance;
		}
		k_test_one_light<true><<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, d.p, h.p, t.p, tape_stride, L.p, u.p);
	} else
		k_test_one_light<false><<<blocks(n, 64), 64, 0, ctx->stream>>>(ctx->scene, n, d.p, h.p, t.p, tape_stride, L.p, u.p);
	NE_FINISH();
	NE_CUDA_OK(L.download(radiance));
	if (used) NE_CUDA_OK(u.download(used));
	return NE_B200_OK;
}

int ne_b200_test_density(ne_b200_ctx* ctx, int n, int instance, const float* ocs_points, float* density, float* inv_max_density) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if ((rc = check_instance(ctx, instance, true))) return rc;
	if (n < 0 || (n && (!ocs_points || !density))) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (inv_max_density) {
		DInstance in; DMaterial m; DVolume v;
		NE_CUDA_OK(cudaMemcpy(&in, ctx->scene.inst + instance, sizeof(in), cudaMemcpyDeviceToHost));
		NE_CUDA_OK(cudaMemcpy(&m, ctx->scene.mat + in.material, sizeof(m), cudaMemcpyDeviceToHost));
		NE_CUDA_OK(cudaMemcpy(&v, ctx->scene.vol + m.volume, sizeof(v), cudaMemcpyDeviceToHost));
		*inv_max_density = v.inv_max_density;
	}
	if (!n) return NE_B200_OK;
	DevBuf<float> p, o;
	NE_UP(p, ocs_points, 3 * size_t(n));
	NE_CUDA_OK(o.alloc(n));
	k_test_density<<<blocks(n, 128), 128, 0, ctx->stream>>>(ctx->scene, n, instance, p.p, o.p);
	NE_FINISH();
	NE_CUDA_OK(o.download(density));
	return NE_B200_OK;
}

int ne_b200_test_read_bricks(ne_b200_ctx* ctx, int volume, int32_t dims[4], int32_t* table, float* inv_majorant, float* pool, float* max_density) {
	int rc = check_ctx(ctx, true);
	if (rc) return rc;
	if (!dims) { set_error("null dims"); return NE_B200_ERR_INVALID; }
	NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	std::vector<DVolume> vols(std::max(1, ctx->nVolumes));
	if (volume < 0 || volume >= ctx->nVolumes) { set_error("volume index out of range"); return NE_B200_ERR_INVALID; }
	NE_CUDA_OK(cudaMemcpy(vols.data(), ctx->scene.vol, sizeof(DVolume) * ctx->nVolumes, cudaMemcpyDeviceToHost));
	const DVolume& v = vols[volume];
	dims[0] = v.bx; dims[1] = v.by; dims[2] = v.bz; dims[3] = v.n_slots;
	if (max_density) *max_density = v.max_density;
	size_t nb = size_t(v.bx) * v.by * v.bz;
	if (table || inv_majorant) {
		std::vector<int2> cells(nb);
		NE_CUDA_OK(cudaMemcpy(cells.data(), v.cells, nb * sizeof(int2), cudaMemcpyDeviceToHost));
		for (size_t b = 0; b < nb; b++) {
			if (table) table[b] = cells[b].x;
			if (inv_majorant) memcpy(&inv_majorant[b], &cells[b].y, 4);
		}
	}
	if (pool && v.n_slots) NE_CUDA_OK(cudaMemcpy(pool, v.pool, size_t(v.n_slots) * BRICK_VOX * sizeof(float), cudaMemcpyDeviceToHost));
	return NE_B200_OK;
}

int ne_b200_test_philox(ne_b200_ctx* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, int n, float* out) {
	int rc = check_ctx(ctx, false);
	if (rc) return rc;
	if (n < 0 || (n && !out)) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	if (!n) return NE_B200_OK;
	DevBuf<float> o;
	NE_CUDA_OK(o.alloc(n));
	k_test_philox<<<blocks(n, 128), 128, 0, ctx->stream>>>(seed, pixel, sample, n, o.p);
	NE_FINISH();
	NE_CUDA_OK(o.download(out));
	return NE_B200_OK;
}

}  // extern "C"
