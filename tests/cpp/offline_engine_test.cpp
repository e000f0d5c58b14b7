// Drives include/ne_b200_offline_engine.hpp the way SceneEditor drives OfflineEngine (SceneEditor.cpp:547-586): a JSON
// scene through ne_b200_scene_file_*, caller-side threads calling renderTile per tile, `pixels` read while tiles land,
// then saveImage-style PNG/EXR export. Prints "OK ..." and exits 0 on success; exits 3 with the library's message when
// no CUDA device is present (there is no CPU fallback). Built and run by tests/test_cpp_adapter.py.
#include <cmath>
#include <cstdio>

#include "ne_b200_offline_engine.hpp"

int main(int argc, char** argv) {
	if (argc < 4) { fprintf(stderr, "usage: %s scene.json resources_dir out_prefix\n", argv[0]); return 2; }
	ne_b200_scene_file* file = nullptr;
	if (ne_b200_scene_file_load(argv[1], argv[2], &file) != NE_B200_OK) { fprintf(stderr, "load: %s\n", ne_b200_last_error()); return 2; }
	ne_b200_camera cam;
	ne_b200_render_settings st;
	ne_b200_scene_file_camera(file, &cam);
	ne_b200_scene_file_settings(file, &st);
	try {
		narval_b200::B200OfflineEngine engine(cam, st, ne_b200_scene_file_desc(file));
		const int W = st.width, H = st.height;
		// one tile in the middle first: only that tile may be filled afterwards
		const int tile = 5 * engine.numberOfTiles.x + 20;
		std::atomic<bool> done{false};
		engine.renderTile(cam, tile, done);
		if (!done) { fprintf(stderr, "tile not flagged finished\n"); return 1; }
		double inTile = 0, outside = 0;
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++) {
				const narval_b200::vec3& p = engine.pixels[size_t(W) * y + x];
				bool in = x / engine.tileSize.x == 20 && y / engine.tileSize.y == 5 && x < 40 * engine.tileSize.x && y < 10 * engine.tileSize.y;
				(in ? inTile : outside) += p.x + p.y + p.z;
			}
		if (!(inTile > 0) || outside != 0) { fprintf(stderr, "tile protocol: in %g outside %g\n", inTile, outside); return 1; }
		engine.coreLoop();  // all tiles from 16 caller-side threads
		double sum = 0;
		for (int y = 0; y < 10 * engine.tileSize.y; y++)
			for (int x = 0; x < 40 * engine.tileSize.x; x++) {
				const narval_b200::vec3& p = engine.pixels[size_t(W) * y + x];
				if (!(p.x >= 0 && p.x <= 1 && p.y >= 0 && p.y <= 1 && p.z >= 0 && p.z <= 1)) { fprintf(stderr, "pixel out of [0,1]\n"); return 1; }
				float lin = engine.linear[3 * (size_t(W) * y + x)];
				float tm = std::pow(1.0f - std::exp(-lin * 0.5f), 1.0f / 2.2f);  // OfflineEngine::postProcessing
				if (std::fabs(tm - p.x) > 1e-5f) { fprintf(stderr, "tone map mismatch %g vs %g\n", tm, p.x); return 1; }
				sum += p.x + p.y + p.z;
			}
		std::string prefix = argv[3];
		if (ne_b200_image_write_png((prefix + ".png").c_str(), W, H, &engine.pixels[0].x) != NE_B200_OK ||
		    ne_b200_image_write_exr((prefix + ".exr").c_str(), W, H, engine.linear.data()) != NE_B200_OK) { fprintf(stderr, "save: %s\n", ne_b200_last_error()); return 1; }
		printf("OK %dx%d spp %d mean %.6f\n", W, H, st.spp, sum / (3.0 * 40 * engine.tileSize.x * 10 * engine.tileSize.y));
	} catch (const std::exception& e) {
		fprintf(stderr, "%s\n", e.what());
		ne_b200_scene_file_free(file);
		return 3;
	}
	ne_b200_scene_file_free(file);
	return 0;
}
