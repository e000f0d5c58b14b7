// c2noise.cpp — WORKLOAD GENERATOR, test infrastructure (see oracle/README.md). Not part of the product.
//
// BASELINE.json configs[1] ("fastnoise-generated 256^3 density") as SURVEY.md §8d C2 specifies it: the reference tree
// vendors FastNoise 0.4.1 (includes/fastnoise/FastNoise.{h,cpp}; unused by the reference's hot path). This file is
// compiled TOGETHER WITH that FastNoise.cpp, in place, by oracle/Makefile into oracle/_ref/libc2noise.so:
//   SimplexFractal (FBM, lacunarity 2, gain 0.5), seed 1337, frequency 4/N, 5 octaves, sampled at voxel indices,
//   value remapped  max(0, n*0.5 + 0.5 - 0.35) / 0.65,  times the radial falloff  smoothstep(0.5, 0.35, |p|)
//   with p = (index + 0.5)/N - 0.5 the voxel centre in the unit cube. Array order [z][y][x], x fastest
//   (Texture(W,H,D,R32F) of ResourceManager::loadVolasTexture, core/ResourceManager.cpp:222-286).
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "FastNoise.h"

extern "C" void c2noise_generate(int W, int H, int D, int seed, float* out) {
	unsigned nt = std::thread::hardware_concurrency();
	if (nt == 0) nt = 1;
	std::vector<std::thread> pool;
	for (unsigned k = 0; k < nt; k++) {
		pool.emplace_back([=]() {
			FastNoise fn(seed);
			fn.SetNoiseType(FastNoise::SimplexFractal);
			fn.SetFractalType(FastNoise::FBM);
			fn.SetFractalOctaves(5);
			fn.SetFrequency(FN_DECIMAL(4.0 / double(W)));
			for (int z = int(k); z < D; z += int(nt))
				for (int y = 0; y < H; y++)
					for (int x = 0; x < W; x++) {
						float n = float(fn.GetSimplexFractal(FN_DECIMAL(x), FN_DECIMAL(y), FN_DECIMAL(z)));
						float v = std::fmax(0.0f, n * 0.5f + 0.5f - 0.35f) / 0.65f;
						float px = (float(x) + 0.5f) / float(W) - 0.5f, py = (float(y) + 0.5f) / float(H) - 0.5f, pz = (float(z) + 0.5f) / float(D) - 0.5f;
						float r = std::sqrt(px * px + py * py + pz * pz);
						float t = (r - 0.5f) / (0.35f - 0.5f);  // smoothstep(edge0 = 0.5, edge1 = 0.35, r)
						t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
						out[(size_t(z) * H + y) * W + x] = v * (t * t * (3.0f - 2.0f * t));
					}
		});
	}
	for (auto& t : pool) t.join();
}
