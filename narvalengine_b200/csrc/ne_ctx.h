// ne_ctx.h — the context object behind ne_b200_ctx and the launch entry points shared by ne_api.cu (C ABI, scene
// upload, test hooks, megakernel) and ne_wavefront.cu (production wavefront renderer).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "ne_b200.h"
#include "ne_scene.cuh"

namespace ne {
void set_error(const std::string& s);
}

#define NE_CUDA_OK(expr)                                                                                      \
	do {                                                                                                      \
		cudaError_t e_ = (expr);                                                                              \
		if (e_ != cudaSuccess) {                                                                              \
			ne::set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                               \
			return NE_B200_ERR_CUDA;                                                                          \
		}                                                                                                     \
	} while (0)

// NE_B200_TRACE_HOST=1: host time of the named scopes on stderr (where an end-to-end frame spends its time outside the kernels)
#include <chrono>
#include <cstdio>
#include <cstdlib>
struct ne_host_span {
	const char* name;
	std::chrono::steady_clock::time_point t0;
	bool on;
	explicit ne_host_span(const char* n) : name(n), on(getenv("NE_B200_TRACE_HOST") != nullptr) {
		if (on) t0 = std::chrono::steady_clock::now();
	}
	~ne_host_span() {
		if (on) fprintf(stderr, "[ne_b200] %-28s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
};

struct ne_wavefront_state;

struct ne_b200_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;     // the stream everything runs on
	cudaStream_t ownStream = nullptr;  // created by ne_b200_create, destroyed with the context
	std::vector<void*> sceneAllocs;  // device allocations owned by the uploaded scene
	ne::DScene scene{};
	bool haveScene = false;
	int nVolumes = 0;
	bool skipWorthwhile = false;  // some volume's brick table is mostly skippable empty space (ne_bricks.cu device_build_majorants)
	int nMeshes = 0;  // triangle meshes with a BVH: the wavefront then runs its persistent trace kernels
	// -1, or the primitive kind (PRIM_RECTANGLE / PRIM_SPHERE / PRIM_POINT) when EVERY light of the scene is a DiffuseLight on
	// that kind and no HomogeneousMedia exists: the shading kernels then run their variant specialised for it
	int lightSet = -1;
	bool onlyGridMedia = false;  // no shadeable surface, no HomogeneousMedia: a camera path that survives its first intersectScene is in a grid medium
	int nSurfaces = 0;  // fold instances a path can be shaded on (a BSDF that is not a medium's): none -> k_wf_surface is never launched
	// world-space corner points (xyz) of everything a camera ray can hit, for the camera-ray culling rectangle
	// (ne_wavefront.cu cull_rect); cullable = false when some instance cannot be bounded or the scene has lights that
	// add radiance to rays that miss everything (directional, environment)
	std::vector<float> boundCorners;
	bool cullable = false;
	unsigned long long pathsCulled = 0;  // camera paths proven to carry no radiance without tracing them (counted as paths)
	unsigned long long sceneGen = 0;  // bumped by every upload: the wavefront's render graph is rebuilt when it changes
	size_t majTableBytes = 0;  // sum of the volumes' 2-byte majorant tables, each padded to 16 bytes (shared-memory staging)
	const void* l2Pool = nullptr;  // the largest brick pool (optional L2 persisting window, NE_B200_L2_PERSIST)
	size_t l2PoolBytes = 0;
	bool testFastShading = false;  // the bsdf / one-light / Li-tape hooks run the FAST medium shading (ne_b200_test_set_fast_shading)
	bool renderPending = false;  // an asynchronous ne_b200_render is in flight: ne_b200_wait checks its outcome
	ne::DCamera cam{};
	bool haveCamera = false;
	float* accum = nullptr;  // W*H*3 fp32 radiance sums
	int W = 0, H = 0;
	int samples = 0;
	ne::DCounters* dCounters = nullptr;
	unsigned long long kernelLaunches = 0, wavefrontIterations = 0;
	// host-side accounts (CUDA events: megakernel, host-driven wavefront loop); the render graph's device-side accounts live
	// in DCounters::stage_ns and are added by ne_b200_get_counters
	double msRender = 0, msVolume = 0, msExtend = 0, msShade = 0, msOther = 0, msUpload = 0;
	cudaEvent_t evA = nullptr, evB = nullptr;
	// the wavefront renderer's state, one per LANE: a render is split into up to two independent sample batches that run on
	// two streams at once, so that one lane's kernels fill the tails of the other's (ne_wavefront.cu, wavefront_render)
	ne_wavefront_state* wf[2] = {nullptr, nullptr};
	cudaStream_t laneStream[2] = {nullptr, nullptr};
	cudaEvent_t laneFork = nullptr, laneJoin[2] = {nullptr, nullptr};
	// device time of asynchronous renders: event pairs recorded around each one on the context's stream, read (and recycled) the
	// next time the counters are fetched
	std::vector<cudaEvent_t> spanEvents;  // pairs: begin, end
	size_t spansPending = 0;               // pairs recorded and not yet folded into msRender
	void* scratch = nullptr;  // reusable device scratch (dense grid staging of the brick builder, resolve buffers)
	size_t scratchBytes = 0;
	void* pinned = nullptr;  // pinned host staging for large uploads from pageable caller memory (ne_bricks.cu h2d_staged)
	size_t pinnedBytes = 0;
};

namespace ne {
// ne_wavefront.cu
int wavefront_render(ne_b200_ctx* ctx, int sppBegin, int sppEnd, int bounces, uint64_t seed, uint32_t flags);
void wavefront_free(ne_b200_ctx* ctx);
size_t wavefront_record_bytes();  // sizeof the wavefront's path + hit record
// ne_bricks.cu
int scratch_reserve(ne_b200_ctx* ctx, size_t bytes);
int device_build_bricks(ne_b200_ctx* ctx, const ne_b200_volume& v, DVolume& out);
int device_build_majorants(ne_b200_ctx* ctx, DVolume& vol);  // DVolume::maj16 / maj_scale from the cells
}  // namespace ne
