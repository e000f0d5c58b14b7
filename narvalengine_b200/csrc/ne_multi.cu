// ne_multi.cu — several B200s of one box behind ONE handle, inside one process: what a single-process host such as
// NarvalEngine's editor (SceneEditor::startOffEngine, src/SceneEditor.cpp:590-604) needs to reach 8 GPUs without writing
// a launcher or a collective itself.
//
//   partition   by SAMPLE INDEX: GPU g of G renders samples [g*spp/G, (g+1)*spp/G) of EVERY pixel on its own
//               scene replica into its own fp32 accumulation buffer. Philox is keyed (seed, pixel, sample), so the union
//               is the set of paths one GPU would trace.
//   execution   one worker thread per device issues that device's work; ne_b200_render is asynchronous (one CUDA graph
//               launch), so all GPUs render concurrently. Scene uploads run on the workers in parallel as well.
//   exchange    ONE kernel on the first device that READS EVERY PEER'S accumulation buffer over NVLink (peer access,
//               vectorised 16-byte loads), sums them in rank order (deterministic), divides by the sample count, applies
//               OfflineEngine::postProcessing and writes the linear and tone-mapped frames: reduce + resolve fused, no
//               intermediate buffer, no second pass. Where peer access is unavailable the buffers are first copied with
//               cudaMemcpyPeerAsync into staging on the first device and the same kernel reads the copies.
// The torchrun arm of bench.py (one PROCESS per GPU, the driver's contract) keeps using an NCCL reduce between processes;
// this file is the single-process path behind the C ABI.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "ne_ctx.h"
#include "ne_device.cuh"

using namespace ne;

namespace {

#define NE_MULTI_MAX 16

struct PeerBuffers {
	const float* p[NE_MULTI_MAX];
	int n;
};

// accum_out (device 0's buffer) <- sum over ranks, in rank order; linear = sum / samples; tonemapped = postProcessing(linear).
// n4 = number of float4 groups (the frame's W*H*3 floats rounded down to a multiple of 4; `tail` floats follow).
__global__ void __launch_bounds__(256) k_multi_reduce_resolve(PeerBuffers in, size_t nFloats, float invSamples, float* accumOut, float* linear,
                                                              float* tonemapped) {
	const size_t n4 = nFloats / 4;
	const size_t stride = size_t(gridDim.x) * blockDim.x;
	for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
		float4 s = __ldg(reinterpret_cast<const float4*>(in.p[0]) + i);
		for (int r = 1; r < in.n; r++) {
			// peer memory over NVLink: streaming loads, every byte is read once
			float4 v = __ldcs(reinterpret_cast<const float4*>(in.p[r]) + i);
			s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
		}
		if (accumOut) reinterpret_cast<float4*>(accumOut)[i] = s;
		float4 m = make_float4(s.x * invSamples, s.y * invSamples, s.z * invSamples, s.w * invSamples);
		if (linear) reinterpret_cast<float4*>(linear)[i] = m;
		if (tonemapped) reinterpret_cast<float4*>(tonemapped)[i] = make_float4(tonemap1(m.x), tonemap1(m.y), tonemap1(m.z), tonemap1(m.w));
	}
	for (size_t i = n4 * 4 + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nFloats; i += stride) {
		float s = in.p[0][i];
		for (int r = 1; r < in.n; r++) s += in.p[r][i];
		if (accumOut) accumOut[i] = s;
		if (linear) linear[i] = s * invSamples;
		if (tonemapped) tonemapped[i] = tonemap1(s * invSamples);
	}
}

// One worker thread per device: runs the jobs posted to it in order (every CUDA call for that device comes from here).
class Worker {
public:
	Worker() : th([this] { loop(); }) {}
	~Worker() {
		{
			std::lock_guard<std::mutex> l(m);
			stop = true;
		}
		cv.notify_all();
		th.join();
	}
	void post(std::function<void()> job) {
		{
			std::lock_guard<std::mutex> l(m);
			jobs.push(std::move(job));
			pending++;
		}
		cv.notify_all();
	}
	void wait() {
		std::unique_lock<std::mutex> l(m);
		idle.wait(l, [this] { return pending == 0; });
	}

private:
	void loop() {
		for (;;) {
			std::function<void()> job;
			{
				std::unique_lock<std::mutex> l(m);
				cv.wait(l, [this] { return stop || !jobs.empty(); });
				if (jobs.empty()) return;
				job = std::move(jobs.front());
				jobs.pop();
			}
			job();
			{
				std::lock_guard<std::mutex> l(m);
				pending--;
			}
			idle.notify_all();
		}
	}
	std::mutex m;
	std::condition_variable cv, idle;
	std::queue<std::function<void()>> jobs;
	int pending = 0;
	bool stop = false;
	std::thread th;
};

}  // namespace

struct ne_b200_multi {
	std::vector<ne_b200_ctx*> ctx;
	std::vector<int> device;
	std::vector<Worker*> worker;
	std::vector<cudaEvent_t> done;    // rank r's render is enqueued up to here
	std::vector<bool> peer;           // device 0 can read rank r's memory directly
	std::vector<float*> staging;      // else: a copy of rank r's buffer on device 0
	size_t stagingFloats = 0;
	std::vector<int> rc;
	std::vector<std::string> err;
	int samples = 0;
};

namespace {

void sample_range(int rank, int world, int spp, int* begin, int* end) {
	const int base = spp / world, rem = spp % world;
	*begin = rank * base + (rank < rem ? rank : rem);
	*end = *begin + base + (rank < rem ? 1 : 0);
}

// Runs fn(rank) on every rank's worker and joins; the first failure is reported with its rank.
int on_all(ne_b200_multi* m, const std::function<int(int)>& fn) {
	const int n = int(m->ctx.size());
	for (int r = 0; r < n; r++) {
		m->rc[r] = NE_B200_OK;
		m->worker[r]->post([m, r, &fn] {
			m->rc[r] = fn(r);
			if (m->rc[r]) m->err[r] = ne_b200_last_error();  // thread-local on the worker: carry it over
		});
	}
	for (int r = 0; r < n; r++) m->worker[r]->wait();
	for (int r = 0; r < n; r++)
		if (m->rc[r]) {
			set_error("GPU " + std::to_string(m->device[r]) + " (rank " + std::to_string(r) + "): " + m->err[r]);
			return m->rc[r];
		}
	return NE_B200_OK;
}

}  // namespace

extern "C" {

int ne_b200_create_multi(const int* gpu_ids, int n_gpus, ne_b200_multi** out) {
	if (!out) { set_error("null out"); return NE_B200_ERR_INVALID; }
	*out = nullptr;
	if (!gpu_ids || n_gpus < 1 || n_gpus > NE_MULTI_MAX) { set_error("bad device list (1.." + std::to_string(NE_MULTI_MAX) + " devices)"); return NE_B200_ERR_INVALID; }
	ne_b200_multi* m = new ne_b200_multi();
	m->rc.assign(n_gpus, 0);
	m->err.assign(n_gpus, "");
	int rc = NE_B200_OK;
	for (int r = 0; r < n_gpus && !rc; r++) {
		ne_b200_ctx* c = nullptr;
		rc = ne_b200_create(gpu_ids[r], &c);
		if (rc) break;
		m->ctx.push_back(c);
		m->device.push_back(gpu_ids[r]);
		m->worker.push_back(new Worker());
		cudaEvent_t e = nullptr;
		if (cudaSetDevice(gpu_ids[r]) != cudaSuccess || cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
			set_error("cudaEventCreate failed");
			rc = NE_B200_ERR_CUDA;
		}
		m->done.push_back(e);
		m->staging.push_back(nullptr);
		// the first device reads the others' accumulation buffers in place when the box allows it (NVLink / NVSwitch)
		bool direct = gpu_ids[r] == gpu_ids[0];
		if (!direct && !rc) {
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, gpu_ids[0], gpu_ids[r]) == cudaSuccess && can) {
				cudaSetDevice(gpu_ids[0]);
				cudaError_t e2 = cudaDeviceEnablePeerAccess(gpu_ids[r], 0);
				direct = e2 == cudaSuccess || e2 == cudaErrorPeerAccessAlreadyEnabled;
				cudaGetLastError();
			}
		}
		m->peer.push_back(direct);
	}
	if (rc) {
		std::string keep = ne_b200_last_error();
		ne_b200_multi_destroy(m);
		set_error(keep);
		return rc;
	}
	*out = m;
	return NE_B200_OK;
}

void ne_b200_multi_destroy(ne_b200_multi* m) {
	if (!m) return;
	for (Worker* w : m->worker) delete w;  // joins
	for (size_t r = 0; r < m->ctx.size(); r++) {
		if (m->staging[r]) { cudaSetDevice(m->device[0]); cudaFree(m->staging[r]); }
		if (m->done[r]) { cudaSetDevice(m->device[r]); cudaEventDestroy(m->done[r]); }
		ne_b200_destroy(m->ctx[r]);
	}
	delete m;
}

int ne_b200_multi_count(const ne_b200_multi* m) { return m ? int(m->ctx.size()) : 0; }

ne_b200_ctx* ne_b200_multi_ctx(ne_b200_multi* m, int rank) {
	if (!m || rank < 0 || rank >= int(m->ctx.size())) return nullptr;
	return m->ctx[rank];
}

int ne_b200_multi_peer_access(const ne_b200_multi* m, int rank) {
	if (!m || rank < 0 || rank >= int(m->ctx.size())) return 0;
	return m->peer[rank] ? 1 : 0;
}

int ne_b200_multi_scene_upload(ne_b200_multi* m, const ne_b200_scene_desc* scene) {
	if (!m || !scene) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	// a replica per GPU, uploaded concurrently (each worker drives its own device's copy engine and brick builder)
	return on_all(m, [m, scene](int r) -> int { return ne_b200_scene_upload(m->ctx[r], scene); });
}

int ne_b200_multi_render(ne_b200_multi* m, const ne_b200_camera* camera, int width, int height, int spp, int bounces, uint64_t seed, uint32_t flags) {
	if (!m) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	if (spp < 0) { set_error("bad render arguments"); return NE_B200_ERR_INVALID; }
	const int n = int(m->ctx.size());
	int rc = on_all(m, [=](int r) -> int {
		ne_b200_ctx* c = m->ctx[r];
		int e;
		if (camera && (e = ne_b200_camera_set(c, camera))) return e;
		if ((e = ne_b200_render(c, width, height, 0, 0, bounces, seed, flags))) return e;  // (re)allocate
		if ((e = ne_b200_clear(c))) return e;
		int b0, b1;
		sample_range(r, n, spp, &b0, &b1);
		if ((e = ne_b200_render(c, width, height, b0, b1, bounces, seed, flags))) return e;  // asynchronous
		NE_CUDA_OK(cudaEventRecord(m->done[r], c->stream));
		return NE_B200_OK;
	});
	if (rc) return rc;
	m->samples = spp;
	return NE_B200_OK;
}

// Joins the render and runs the fused reduce + resolve on the first device. Host pointers (either may be NULL).
int ne_b200_multi_resolve(ne_b200_multi* m, float* pixels_tonemapped, float* pixels_linear) {
	if (!m) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	const int n = int(m->ctx.size());
	ne_b200_ctx* c0 = m->ctx[0];
	if (!c0->accum || m->samples <= 0) { set_error("nothing rendered yet"); return NE_B200_ERR_STATE; }
	NE_CUDA_OK(cudaSetDevice(m->device[0]));
	const size_t nFloats = size_t(c0->W) * c0->H * 3;
	PeerBuffers in;
	in.n = n;
	for (int r = 0; r < n; r++) {
		ne_b200_ctx* c = m->ctx[r];
		if (c->W != c0->W || c->H != c0->H || !c->accum) { set_error("ranks disagree on the frame"); return NE_B200_ERR_STATE; }
		if (r > 0) NE_CUDA_OK(cudaStreamWaitEvent(c0->stream, m->done[r], 0));  // rank r's render, on its own device's stream
		in.p[r] = c->accum;
		if (!m->peer[r]) {
			if (m->stagingFloats < nFloats) {
				for (float*& s : m->staging) { if (s) cudaFree(s); s = nullptr; }
				m->stagingFloats = nFloats;
			}
			if (!m->staging[r]) NE_CUDA_OK(cudaMalloc(&m->staging[r], nFloats * sizeof(float)));
			NE_CUDA_OK(cudaMemcpyPeerAsync(m->staging[r], m->device[0], c->accum, m->device[r], nFloats * sizeof(float), c0->stream));
			in.p[r] = m->staging[r];
		}
	}
	int rc;
	if ((rc = scratch_reserve(c0, 2 * nFloats * sizeof(float)))) return rc;
	float* lin = static_cast<float*>(c0->scratch);
	float* tm = lin + nFloats;
	cudaDeviceProp prop;
	NE_CUDA_OK(cudaGetDeviceProperties(&prop, m->device[0]));
	const int grid = prop.multiProcessorCount * 8;
	// the sum lands in rank 0's accumulation buffer too: checkpoints (ne_b200_accum_download on rank 0) see the whole frame
	k_multi_reduce_resolve<<<grid, 256, 0, c0->stream>>>(in, nFloats, 1.0f / float(m->samples), c0->accum, pixels_linear ? lin : nullptr,
	                                                      pixels_tonemapped ? tm : nullptr);
	c0->kernelLaunches++;
	NE_CUDA_OK(cudaGetLastError());
	if (pixels_linear) NE_CUDA_OK(cudaMemcpyAsync(pixels_linear, lin, nFloats * sizeof(float), cudaMemcpyDeviceToHost, c0->stream));
	if (pixels_tonemapped) NE_CUDA_OK(cudaMemcpyAsync(pixels_tonemapped, tm, nFloats * sizeof(float), cudaMemcpyDeviceToHost, c0->stream));
	NE_CUDA_OK(cudaStreamSynchronize(c0->stream));
	c0->samples = m->samples;
	// outcome of every rank's render (overflow flag, sticky errors)
	return on_all(m, [m](int r) -> int { return ne_b200_wait(m->ctx[r]); });
}

int ne_b200_multi_render_frame(ne_b200_multi* m, const ne_b200_camera* camera, int width, int height, int spp, int bounces, uint64_t seed, uint32_t flags,
                               float* pixels_tonemapped, float* pixels_linear) {
	int rc = ne_b200_multi_render(m, camera, width, height, spp, bounces, seed, flags);
	if (rc) return rc;
	return ne_b200_multi_resolve(m, pixels_tonemapped, pixels_linear);
}

}  // extern "C"
