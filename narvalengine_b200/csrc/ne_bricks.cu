// ne_bricks.cu — the brick-sparse density grid built ON THE GPU from a dense W*H*D grid (the Texture that
// ResourceManager::loadVolasTexture / loadVDBasTexture hand to GridMedia, core/ResourceManager.cpp:165-286).
// ne_b200_scene_upload copies the caller's dense grid to HBM once and three kernels turn it into the layout the
// tracking kernels walk (ne_scene.cuh DVolume): per brick {slot, 1/majorant} cells, 9^3 apron records, global maximum.
// The result is bit-identical to the host builder ne_b200_host_build_bricks (ne_host.cpp), which stays the
// inspectable definition and serves leaf (.vdb) input; tests/test_gpu_parity.py compares the two.
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_fp16.h>

#include "ne_ctx.h"

using namespace ne;

namespace {

// One warp per brick: need = any voxel != 0 in the record's [8b, 8b+8]^3; maj = max(0, voxels of [8b-1, 8b+8]^3)
// (the trilinear stencil of any point of the brick, +-1 voxel); voxels at or beyond the grid count as 0
// (GridMedia::density :17-18). Also the grid's global maximum (GridMedia::calculateMaxDensity, GridMedia.h:16-21).
__global__ void __launch_bounds__(256) k_brick_scan(const float* __restrict__ dense, int W, int H, int D, int nbx, int nby, int nbz,
                                                    uint32_t* __restrict__ need, float* __restrict__ maj, unsigned int* gmaxBits) {
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	const int nb = nbx * nby * nbz;
	if (warp >= nb) return;
	const int bx = warp % nbx, by = (warp / nbx) % nby, bz = warp / (nbx * nby);
	bool any = false;
	float m = 0.0f;
	for (int k = lane; k < 1000; k += 32) {  // the 10^3 voxels [8b-1, 8b+8]^3
		int lx = k % 10 - 1, ly = (k / 10) % 10 - 1, lz = k / 100 - 1;
		int x = bx * 8 + lx, y = by * 8 + ly, z = bz * 8 + lz;
		if (x < 0 || y < 0 || z < 0 || x >= W || y >= H || z >= D) continue;
		float v = dense[(size_t(z) * H + y) * W + x];
		m = v > m ? v : m;  // std::max(m, v) of the host builder
		if (lx >= 0 && ly >= 0 && lz >= 0 && v != 0.0f) any = true;
	}
	for (int o = 16; o; o >>= 1) {
		float t = __shfl_xor_sync(0xffffffffu, m, o);
		m = t > m ? t : m;
	}
	any = __any_sync(0xffffffffu, any);
	if (lane == 0) {
		need[warp] = any ? 1u : 0u;
		maj[warp] = m;
		if (m > 0) atomicMax(gmaxBits, __float_as_uint(m));  // positive floats order like their bit patterns
	}
}

// Exclusive prefix sum of need[] (raster order over bricks = the host builder's slot order) by one block.
__global__ void __launch_bounds__(1024) k_brick_slots(const uint32_t* __restrict__ need, int nb, int* __restrict__ slot, int* total) {
	__shared__ uint32_t warpSum[32];
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int base = 0; base < nb; base += 1024) {
		int i = base + threadIdx.x;
		uint32_t v = i < nb ? need[i] : 0u, x = v;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o) x += t;
		}
		if (lane == 31) warpSum[w] = x;
		__syncthreads();
		if (w == 0) {
			uint32_t s = warpSum[lane], y = s;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t t = __shfl_up_sync(0xffffffffu, y, o);
				if (lane >= o) y += t;
			}
			warpSum[lane] = y - s;  // exclusive over warps
		}
		__syncthreads();
		uint32_t excl = carry + warpSum[w] + x - v;
		if (i < nb) slot[i] = v ? int(excl) : -1;
		__syncthreads();
		if (threadIdx.x == 1023) carry = excl + v;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = int(carry);
}

// One block per brick: the {slot, 1/majorant} cell and, for bricks with storage, the 9^3 apron record.
__global__ void __launch_bounds__(256) k_brick_fill(const float* __restrict__ dense, int W, int H, int D, int nbx, int nby, const int* __restrict__ slot,
                                                    const float* __restrict__ maj, int2* __restrict__ cells, float* __restrict__ pool) {
	const int b = blockIdx.x;
	const int s = slot[b];
	if (threadIdx.x == 0) {
		float m = maj[b];
		float inv = (s >= 0 && m > 0) ? 1.0f / m : 0.0f;
		cells[b] = make_int2(s, __float_as_int(inv));
	}
	if (s < 0) return;
	const int bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
	float* dst = pool + size_t(s) * BRICK_VOX;
	for (int k = threadIdx.x; k < BRICK_VOX; k += blockDim.x) {
		int x = bx * 8 + k % 9, y = by * 8 + (k / 9) % 9, z = bz * 8 + k / 81;
		dst[k] = (x < W && y < H && z < D) ? dense[(size_t(z) * H + y) * W + x] : 0.0f;
	}
}

}  // namespace

namespace ne {

int scratch_reserve(ne_b200_ctx* ctx, size_t bytes) {
	if (ctx->scratchBytes >= bytes) return NE_B200_OK;
	if (ctx->scratch) {
		NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->scratch);
		ctx->scratch = nullptr;
		ctx->scratchBytes = 0;
	}
	NE_CUDA_OK(cudaMalloc(&ctx->scratch, bytes));
	ctx->scratchBytes = bytes;
	return NE_B200_OK;
}

// Host -> device copy of a large PAGEABLE caller buffer. cudaMemcpy from pageable memory is staged by the driver on one
// thread (~7 GB/s measured for the 64 MiB C2 grid = 9 ms of a 49 ms end-to-end frame). Here a few host threads copy
// 4 MiB chunks into the context's pinned staging buffer while the calling thread issues one async DMA per chunk as it
// becomes ready, so the memcpy runs at several threads' worth of memory bandwidth and overlaps the PCIe transfer.
// Buffers larger than the staging cap go in rounds.
int h2d_staged(ne_b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
	const size_t CH = size_t(4) << 20, CAP = size_t(256) << 20;
	// a caller that already holds the data in page-locked memory (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch
	// tensor) gets one direct DMA: no staging copy at all
	cudaPointerAttributes attr{};
	const bool srcPinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();  // an unregistered host pointer is not an error here
	if (bytes < 2 * CH || srcPinned) {
		NE_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
		return NE_B200_OK;
	}
	const size_t want = std::min(bytes, CAP);
	if (ctx->pinnedBytes < want) {
		if (ctx->pinned) { NE_CUDA_OK(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; ctx->pinnedBytes = 0; }
		NE_CUDA_OK(cudaHostAlloc(&ctx->pinned, want, cudaHostAllocDefault));
		ctx->pinnedBytes = want;
	}
	const unsigned hw = std::thread::hardware_concurrency();
	const int nThreads = int(std::max(1u, std::min(8u, hw ? hw / 2 : 4u)));
	for (size_t roundOff = 0; roundOff < bytes; roundOff += ctx->pinnedBytes) {
		const size_t roundBytes = std::min(ctx->pinnedBytes, bytes - roundOff);
		const size_t nChunks = (roundBytes + CH - 1) / CH;
		std::vector<std::atomic<int>> ready(nChunks);
		for (auto& r : ready) r.store(0, std::memory_order_relaxed);
		std::atomic<size_t> next{0};
		const char* s8 = static_cast<const char*>(src) + roundOff;
		char* pin = static_cast<char*>(ctx->pinned);
		auto worker = [&]() {
			for (size_t c = next++; c < nChunks; c = next++) {
				size_t off = c * CH, len = std::min(CH, roundBytes - off);
				memcpy(pin + off, s8 + off, len);
				ready[c].store(1, std::memory_order_release);
			}
		};
		std::vector<std::thread> pool;
		for (int t = 0; t < nThreads; t++) pool.emplace_back(worker);
		cudaError_t err = cudaSuccess;
		for (size_t c = 0; c < nChunks; c++) {
			while (!ready[c].load(std::memory_order_acquire)) std::this_thread::yield();
			size_t off = c * CH, len = std::min(CH, roundBytes - off);
			if (err == cudaSuccess) err = cudaMemcpyAsync(static_cast<char*>(dst) + roundOff + off, pin + off, len, cudaMemcpyHostToDevice, ctx->stream);
		}
		for (auto& t : pool) t.join();
		NE_CUDA_OK(err);
		// the staging buffer is reused by the next round / the next upload
		NE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	}
	return NE_B200_OK;
}

// dense (host) -> DVolume in HBM. Allocations that belong to the scene are pushed to ctx->sceneAllocs.
int device_build_bricks(ne_b200_ctx* ctx, const ne_b200_volume& v, DVolume& out) {
	ne_host_span span_("  device_build_bricks");
	const size_t nvox = size_t(v.width) * v.height * v.depth;
	const int nbx = (v.width + 7) / 8, nby = (v.height + 7) / 8, nbz = (v.depth + 7) / 8;
	const size_t nb = size_t(nbx) * nby * nbz;
	if (nb > size_t(1) << 30) { set_error("grid too large"); return NE_B200_ERR_INVALID; }
	// scratch: dense grid | need | maj | slot | {total, gmax}
	auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
	const size_t oDense = 0, oNeed = up(nvox * 4), oMaj = oNeed + up(nb * 4), oSlot = oMaj + up(nb * 4), oTot = oSlot + up(nb * 4);
	int rc = scratch_reserve(ctx, oTot + 256);
	if (rc) return rc;
	char* base = static_cast<char*>(ctx->scratch);
	float* dDense = reinterpret_cast<float*>(base + oDense);
	uint32_t* dNeed = reinterpret_cast<uint32_t*>(base + oNeed);
	float* dMaj = reinterpret_cast<float*>(base + oMaj);
	int* dSlot = reinterpret_cast<int*>(base + oSlot);
	int* dTot = reinterpret_cast<int*>(base + oTot);
	cudaStream_t st = ctx->stream;
	NE_CUDA_OK(cudaMemsetAsync(dTot, 0, 8, st));
	if ((rc = h2d_staged(ctx, dDense, v.dense, nvox * 4))) return rc;
	k_brick_scan<<<unsigned((nb * 32 + 255) / 256), 256, 0, st>>>(dDense, v.width, v.height, v.depth, nbx, nby, nbz, dNeed, dMaj,
	                                                             reinterpret_cast<unsigned int*>(dTot + 1));
	k_brick_slots<<<1, 1024, 0, st>>>(dNeed, int(nb), dSlot, dTot);
	ctx->kernelLaunches += 2;
	int tot[2] = {0, 0};
	NE_CUDA_OK(cudaMemcpyAsync(tot, dTot, 8, cudaMemcpyDeviceToHost, st));
	NE_CUDA_OK(cudaStreamSynchronize(st));
	int2* dCells = nullptr;
	float* dPool = nullptr;
	NE_CUDA_OK(cudaMallocAsync(&dCells, nb * sizeof(int2), st));
	ctx->sceneAllocs.push_back(dCells);
	NE_CUDA_OK(cudaMallocAsync(&dPool, std::max<size_t>(1, size_t(tot[0]) * BRICK_VOX) * sizeof(float), st));
	ctx->sceneAllocs.push_back(dPool);
	k_brick_fill<<<unsigned(nb), 256, 0, st>>>(dDense, v.width, v.height, v.depth, nbx, nby, dSlot, dMaj, dCells, dPool);
	ctx->kernelLaunches++;
	NE_CUDA_OK(cudaGetLastError());
	out.W = v.width; out.H = v.height; out.D = v.depth;
	out.bx = nbx; out.by = nby; out.bz = nbz;
	out.n_slots = tot[0];
	out.cells = dCells;
	out.pool = dPool;
	memcpy(&out.max_density, &tot[1], 4);
	out.inv_max_density = 1.0f / out.max_density;
	return NE_B200_OK;
}

// The 2-byte table the tracking walks read at every brick crossing.
//   brick with a record:  q in [1, 32767] = its majorant as a multiple of maj_scale, rounded UP (the walk's majorant
//                         q * maj_scale bounds the brick's maximum from above: tracking stays unbiased)
//   empty brick:          0x8000 | d, d = Chebyshev distance (in bricks, capped at NE_SKIP_MAX) to the nearest brick with
//                         a record OR to the outside of the table: every brick closer than d is empty and inside, so a
//                         walk standing here may cross the whole (2d-1)^3 cube in one move (BrickTracker::jump)
#define NE_SKIP_MAX 16
static __global__ void k_brick_maj16(const int2* __restrict__ cells, int n, float scale, unsigned short* __restrict__ out) {
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n) return;
	float inv = __int_as_float(cells[b].y);
	unsigned q = 0x8001u;
	if (cells[b].x >= 0 && inv > 0) {
		float m = 1.0f / inv;
		q = (unsigned)ceilf(m / scale);
		if (q < 1) q = 1;
		while (q < 32767u && float(q) * scale < m) q++;
		if (q > 32767u) q = 32767u;
	}
	out[b] = (unsigned short)q;
}
// One relaxation of the distance field: d_i(c) = min(NE_SKIP_MAX, 1 + min over the 26 neighbours of d_{i-1}), with d = 0
// for bricks with a record and for the outside. Started from d_0 = 1, i iterations give min(true distance, i + 1).
static __global__ void k_brick_skip(const unsigned short* __restrict__ in, unsigned short* __restrict__ out, int nbx, int nby, int nbz) {
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nbx * nby * nbz) return;
	unsigned short v = in[b];
	if (!(v & 0x8000u)) { out[b] = v; return; }
	int x = b % nbx, y = (b / nbx) % nby, z = b / (nbx * nby);
	unsigned best = NE_SKIP_MAX;
	for (int dz = -1; dz <= 1; dz++)
		for (int dy = -1; dy <= 1; dy++)
			for (int dx = -1; dx <= 1; dx++) {
				int xx = x + dx, yy = y + dy, zz = z + dz;
				unsigned d = 0;
				if (unsigned(xx) < unsigned(nbx) && unsigned(yy) < unsigned(nby) && unsigned(zz) < unsigned(nbz)) {
					unsigned short w = in[(zz * nby + yy) * nbx + xx];
					d = (w & 0x8000u) ? (w & 0x7fffu) : 0u;
				}
				best = min(best, d);
			}
	out[b] = (unsigned short)(0x8000u | min(unsigned(NE_SKIP_MAX), best + 1));
}

// Short skips cost more than the bricks they save (the jump is ~40 instructions and the warp's other lanes wait for
// it): distances below `minD` are stored as 1 = "no skip".
static __global__ void k_brick_skip_floor(unsigned short* __restrict__ t, int n, unsigned minD, unsigned* __restrict__ nSkippable) {
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	bool far = false;
	if (b < n) {
		unsigned short v = t[b];
		far = (v & 0x8000u) && (v & 0x7fffu) >= minD;
		if ((v & 0x8000u) && !far) t[b] = (unsigned short)0x8001u;
	}
	unsigned m = __ballot_sync(0xffffffffu, far);
	if ((threadIdx.x & 31) == 0 && m) atomicAdd(nSkippable, (unsigned)__popc(m));
}

// The table the walks read (layout and encoding: ne_tracking.cuh, BrickTracker): the bricks with a one-brick apron, IEEE
// halves. codes[] = the un-aproned 16-bit codes above (only "has a record" vs "empty, distance d" is taken from them; the
// majorant itself comes from the cell, times K = 2^k, rounded UP to the next half).
static __global__ void k_brick_table(const int2* __restrict__ cells, const unsigned short* __restrict__ codes, int nbx, int nby, int nbz, float K,
                                     unsigned short* __restrict__ out) {
	const int SY = nbx + 2, SZ = SY * (nby + 2);
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= SZ * (nbz + 2)) return;
	const int x = i % SY - 1, y = (i / SY) % (nby + 2) - 1, z = i / SZ - 1;
	unsigned short h = 0xFC00u;  // -inf: outside the grid
	if (unsigned(x) < unsigned(nbx) && unsigned(y) < unsigned(nby) && unsigned(z) < unsigned(nbz)) {
		const int b = (z * nby + y) * nbx + x;
		const unsigned short c = codes[b];
		const float inv = __int_as_float(cells[b].y);
		if (!(c & 0x8000u) && cells[b].x >= 0 && inv > 0) {
			h = __half_as_ushort(__float2half_ru((1.0f / inv) * K));
			if (h == 0) h = 1;             // below the smallest subnormal half: still an upper bound
			if (h >= 0x7C00u) h = 0x7BFFu;  // cannot happen: K puts the global maximum below 2^14
		} else {
			const float d = float(max(1u, unsigned(c & 0x7fffu)));
			h = __half_as_ushort(__float2half_rn(-d));  // 1..16: exact
		}
	}
	out[i] = h;
}

int device_build_majorants(ne_b200_ctx* ctx, DVolume& vol) {
	ne_host_span span_("  device_build_majorants");
	const int nb = vol.bx * vol.by * vol.bz;
	const int nt = (vol.bx + 2) * (vol.by + 2) * (vol.bz + 2);
	cudaStream_t st = ctx->stream;
	unsigned short *d = nullptr, *tmp = nullptr, *table = nullptr;
	NE_CUDA_OK(cudaMallocAsync(&d, size_t(std::max(1, nb)) * sizeof(unsigned short), st));
	// padded to 16 bytes: the tracking kernels copy the table to shared memory with bulk async copies (16-byte granules)
	NE_CUDA_OK(cudaMallocAsync(&table, ((size_t(nt) * sizeof(unsigned short) + 15) & ~size_t(15)) + 16, st));
	ctx->sceneAllocs.push_back(table);
	// the scale of the intermediate 16-bit codes (only their empty / non-empty split and the distance field are used)
	const float codeScale = vol.max_density > 0 ? vol.max_density * (1.0f / 32767.0f) * 1.000001f : 1.0f;
	// majorant * 2^k as a half: k puts the global maximum into [2^13, 2^14), far from the half range's ends
	int e = 0;
	if (vol.max_density > 0 && std::isfinite(vol.max_density)) frexpf(vol.max_density, &e);
	const float K = ldexpf(1.0f, 14 - e);
	vol.maj_scale = ldexpf(1.0f, e - 14);
	if (nb > 0) {
		NE_CUDA_OK(cudaMallocAsync(&tmp, nb * sizeof(unsigned short), st));
		const unsigned grid = unsigned((nb + 255) / 256);
		// an odd number of relaxations, so the last one lands in `d`
		k_brick_maj16<<<grid, 256, 0, st>>>(vol.cells, nb, codeScale, tmp);
		unsigned short *src = tmp, *dst = d;
		for (int i = 0; i < NE_SKIP_MAX - 1; i++) {
			k_brick_skip<<<grid, 256, 0, st>>>(src, dst, vol.bx, vol.by, vol.bz);
			std::swap(src, dst);
		}
		const char* e = getenv("NE_B200_SKIP_MIN");  // smallest cube radius worth a jump (>= NE_SKIP_MAX: no skipping at all)
		const unsigned minR = e ? unsigned(strtoul(e, nullptr, 10)) : 2u;
		unsigned* dCount = nullptr;
		NE_CUDA_OK(cudaMallocAsync(&dCount, sizeof(unsigned), st));
		NE_CUDA_OK(cudaMemsetAsync(dCount, 0, sizeof(unsigned), st));
		k_brick_skip_floor<<<grid, 256, 0, st>>>(d, nb, minR + 1, dCount);
		ctx->kernelLaunches += NE_SKIP_MAX + 1;
		NE_CUDA_OK(cudaGetLastError());
		unsigned nSkippable = 0;
		NE_CUDA_OK(cudaMemcpyAsync(&nSkippable, dCount, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
		NE_CUDA_OK(cudaStreamSynchronize(st));
		NE_CUDA_OK(cudaFreeAsync(dCount, st));
		NE_CUDA_OK(cudaFreeAsync(tmp, st));
		// the skipping variant of the tracking kernels costs ~4 % where there is little to skip (C2: 17.3 vs 16.6 ms) and
		// saves 20-25 % on a WDAS-scale sparse cloud: use it when a third of the table is far from any density
		if (double(nSkippable) >= 0.33 * double(nb)) ctx->skipWorthwhile = true;
	}
	k_brick_table<<<unsigned((nt + 255) / 256), 256, 0, st>>>(vol.cells, d, vol.bx, vol.by, vol.bz, K, table);
	ctx->kernelLaunches++;
	NE_CUDA_OK(cudaGetLastError());
	NE_CUDA_OK(cudaFreeAsync(d, st));
	vol.maj16 = table;
	return NE_B200_OK;
}

}  // namespace ne
