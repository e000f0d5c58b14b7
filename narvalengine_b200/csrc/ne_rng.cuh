// ne_rng.cuh — uniform sources for the estimator.
//   TapeRng    replays an explicit uniform tape in the reference's draw order: what
//              narvalengine::random() (src/utils/Math.h:59-66) returned after mt.seed(k). Test hooks only.
//   PhiloxRng  Philox4x32-10 (Salmon et al., SC'11), counter-based: key = 64-bit seed, counter =
//              (pixel, sample, dimension/4, stream). Any (pixel, sample) stream is reproducible on any GPU in any order.
#pragma once
#include "ne_math.cuh"

namespace ne {

struct TapeRng {
	const float* tape;
	int n, pos;
	bool overflow;
	NE_D void init(const float* t, int len) { tape = t; n = len; pos = 0; overflow = false; }
	NE_D float next() {
		if (pos >= n) { overflow = true; return 0.5f; }
		return tape[pos++];
	}
	NE_D void begin_event() {}
};

NE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
	const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
	for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
		uint32_t hi0 = __umulhi(M0, c0), hi1 = __umulhi(M1, c2);
#else
		uint32_t hi0 = uint32_t((uint64_t(M0) * c0) >> 32), hi1 = uint32_t((uint64_t(M1) * c2) >> 32);
#endif
		uint32_t lo0 = M0 * c0, lo1 = M1 * c2;
		uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += W0; k1 += W1;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

NE_HD float u32_to_unit(uint32_t x) { return float(x >> 8) * (1.0f / 16777216.0f); }  // [0, 1 - 2^-24]

struct PhiloxRng {
	uint32_t k0, k1, pixel, sample, dim, stream;
	uint32_t b0, b1, b2, b3;
	bool fresh;  // the block of the current (4-aligned) dimension is already in b0..b3
	NE_D void init(uint64_t seed, uint32_t px, uint32_t smp, uint32_t dimension = 0, uint32_t strm = 0) {
		k0 = uint32_t(seed); k1 = uint32_t(seed >> 32); pixel = px; sample = smp; dim = dimension; stream = strm;
		fresh = false;
		if (dim & 3) refill();
	}
	// Start a tracking event on a fresh block: all lanes of a warp generate their block HERE, together, instead of
	// each lane refilling inside whichever next() happens to cross a multiple of four (a divergent ~70-instruction branch).
	NE_D void begin_event() {
		dim = (dim + 3u) & ~3u;
		refill();
		fresh = true;
	}
	NE_D void refill() {
		uint32_t o[4];
		philox4x32_10(pixel, sample, dim >> 2, stream, k0, k1, o);
		b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
	}
	NE_D float next() {
		uint32_t l = dim & 3;
		if (l == 0 && !fresh) refill();
		fresh = false;
		dim++;
		uint32_t x = l == 0 ? b0 : (l == 1 ? b1 : (l == 2 ? b2 : b3));
		return u32_to_unit(x);
	}
};

// A side stream for work that is evaluated out of line (the transmittance walk of a next-event request runs in
// its own wavefront kernel): Philox forks to counter word 3 = `stream`, dimension 0; a tape just continues.
template <class R>
struct Fork {
	R& r;
	NE_D Fork(R& base, uint32_t) : r(base) {}
	NE_D R& get() { return r; }
};
template <>
struct Fork<PhiloxRng> {
	PhiloxRng f;
	NE_D Fork(PhiloxRng& base, uint32_t stream) {
		f = base;
		f.stream = stream;
		f.dim = 0;
		f.fresh = false;
	}
	NE_D PhiloxRng& get() { return f; }
};

}  // namespace ne
