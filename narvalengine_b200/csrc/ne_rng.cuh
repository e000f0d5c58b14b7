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
	NE_D void next_block(uint32_t w[4]) {
		for (int k = 0; k < 4; k++) w[k] = uint32_t(next() * 16777216.0f) * 2654435761u;
	}
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
	NE_D void init(uint64_t seed, uint32_t px, uint32_t smp, uint32_t dimension = 0, uint32_t strm = 0) {
		k0 = uint32_t(seed); k1 = uint32_t(seed >> 32); pixel = px; sample = smp; dim = dimension; stream = strm;
		if (dim & 3) refill();
	}
	NE_D void refill() {
		uint32_t o[4];
		philox4x32_10(pixel, sample, dim >> 2, stream, k0, k1, o);
		b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
	}
	NE_D float next() {
		uint32_t l = dim & 3;
		if (l == 0) refill();
		dim++;
		uint32_t x = l == 0 ? b0 : (l == 1 ? b1 : (l == 2 ? b2 : b3));
		return u32_to_unit(x);
	}
	// One whole block (128 bits) at the next 4-aligned dimension: the key of a tracking walk's own stream (PcgRng).
	NE_D void next_block(uint32_t w[4]) {
		dim = (dim + 3u) & ~3u;
		philox4x32_10(pixel, sample, dim >> 2, stream, k0, k1, w);
		dim += 4;
	}
};

// The uniform stream of ONE tracking walk in production (per-brick-majorant) mode: PCG32 (XSH-RR 64/32, O'Neill 2014)
// whose 64-bit state and stream selector are one Philox block of the path's (seed, pixel, sample, dimension) counter.
// Every walk is therefore still a pure function of the Philox key - reproducible on any GPU in any order - but a
// free-flight/acceptance draw costs ~10 instructions instead of a share of a 10-round Philox block generated at a
// divergent point of the event loop. A walk that is cut (budget) and resumed re-keys from the next Philox block.
struct PcgRng {
	uint64_t state, inc;
	template <class R>
	NE_D void start(R& base) {
		uint32_t w[4];
		base.next_block(w);
		state = (uint64_t(w[0]) << 32) | w[1];
		inc = ((uint64_t(w[2]) << 32) | w[3]) | 1ull;
		state = state * 6364136223846793005ull + inc;
	}
	NE_D float next() {
		uint64_t old = state;
		state = old * 6364136223846793005ull + inc;
		uint32_t xs = uint32_t(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = uint32_t(old >> 59u);
		return u32_to_unit((xs >> rot) | (xs << ((32u - rot) & 31u)));
	}
};
// Reference-order walks (global majorant, tape tests) draw straight from the path's own source.
template <class R>
struct PassRng {
	R* r;
	NE_D void start(R& base) { r = &base; }
	NE_D float next() { return r->next(); }
};
template <bool BRICKMAJ, class R>
struct WalkRngOf { typedef PassRng<R> type; };
template <class R>
struct WalkRngOf<true, R> { typedef PcgRng type; };

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
	}
	NE_D PhiloxRng& get() { return f; }
};

}  // namespace ne
