// ne_b200_offline_engine.hpp — the renderer seam of NarvalEngine (OfflineEngine, src/core/OfflineEngine.h:23-47,
// .cpp:4-139) on top of the C ABI, as a header-only C++ class with no dependency on the engine's own headers.
//
// It keeps OfflineEngine's public surface and tile protocol: the caller owns the threads, calls
// renderTile(camera, index, finished) once per tile (row-major over numberOfTiles, mx = index % numberOfTiles.x,
// SceneEditor.cpp:547-586), may read `pixels` (tone-mapped, glm::vec3-compatible, row-major W*y+x) at any time, and
// sees `finished = true` when tile `index` is final. The FIRST renderTile of a frame renders the whole frame on the GPU
// (one ne_b200_render_frame); every renderTile(i) then copies its tile into `pixels`. updateOfflineEngine() starts a
// new frame. Inside NarvalEngine the only extra code is the scene flattening shown in INTEGRATION.md §2 (Scene* ->
// ne_b200_scene_desc) and `toPod(Camera)`; here both arrive as the ABI's POD types, e.g. from ne_b200_scene_file_*.
#pragma once
#include <atomic>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ne_b200.h"

namespace narval_b200 {

struct vec3 { float x, y, z; };    // layout of glm::vec3 (OfflineEngine::pixels)
struct ivec2 { int x, y; };

class B200OfflineEngine {
public:
	// OfflineEngine.h:25-38
	ne_b200_camera camera{};
	ne_b200_render_settings settings{};
	int numberOfThreads = 16;
	ivec2 numberOfTiles{40, 10}, tileSize{0, 0};
	std::thread threadPool[16];
	std::atomic<bool> isThreadDone[16];
	vec3* pixels = nullptr;           // tone-mapped frame, what SceneEditor uploads as RGB32F (SceneEditor.cpp:560)
	std::vector<float> linear;        // un-tone-mapped frame (color / spp), for EXR export
	uint64_t seed = 1;
	uint32_t flags = 0;

	B200OfflineEngine(const ne_b200_camera& cam, const ne_b200_render_settings& st, const ne_b200_scene_desc* scene, int cudaDevice = 0)
	    : B200OfflineEngine(cam, st, scene, std::vector<int>{cudaDevice}) {}
	// Several GPUs of the box: scene replicas, the frame's samples split by index, one fused reduce + resolve kernel over
	// peer memory on the first device (ne_b200_create_multi). One device = the single-GPU path.
	B200OfflineEngine(const ne_b200_camera& cam, const ne_b200_render_settings& st, const ne_b200_scene_desc* scene, const std::vector<int>& cudaDevices) {
		for (auto& d : isThreadDone) d = false;
		check(ne_b200_create_multi(cudaDevices.data(), int(cudaDevices.size()), &multi), "ne_b200_create_multi");
		ctx = ne_b200_multi_ctx(multi, 0);
		updateOfflineEngine(cam, st, scene);
	}
	~B200OfflineEngine() {
		delete[] pixels;  // the engine owns `pixels` (OfflineEngine.cpp:4-6), not the scene or the camera
		ne_b200_multi_destroy(multi);
	}
	B200OfflineEngine(const B200OfflineEngine&) = delete;
	B200OfflineEngine& operator=(const B200OfflineEngine&) = delete;

	// OfflineEngine::updateOfflineEngine(Camera, SceneSettings, Scene*): new camera / settings / scene, new frame.
	void updateOfflineEngine(const ne_b200_camera& cam, const ne_b200_render_settings& st, const ne_b200_scene_desc* scene) {
		std::lock_guard<std::mutex> lock(frameMutex);
		camera = cam;
		settings = st;
		tileSize = ivec2{settings.width / numberOfTiles.x, settings.height / numberOfTiles.y};  // integer division (Q28)
		delete[] pixels;
		const size_t n = size_t(settings.width) * settings.height;
		pixels = new vec3[n]();       // readable (zeros) before any tile is final
		frame.assign(n * 3, 0.0f);
		linear.assign(n * 3, 0.0f);
		if (scene) check(ne_b200_multi_scene_upload(multi, scene), "ne_b200_multi_scene_upload");
		frameReady = false;
	}

	// OfflineEngine::renderTile(Camera, int, std::atomic<bool>&), OfflineEngine.cpp:54-76
	void renderTile(const ne_b200_camera& cam, int index, std::atomic<bool>& finished) {
		{
			std::lock_guard<std::mutex> lock(frameMutex);
			if (!frameReady) {  // the first tile of a frame kicks the whole frame on the GPU
				check(ne_b200_multi_render_frame(multi, &cam, settings.width, settings.height, settings.spp, settings.bounces, seed, flags, frame.data(),
				                                 linear.data()), "ne_b200_multi_render_frame");
				frameReady = true;
			}
		}
		const int mx = index % numberOfTiles.x, my = index / numberOfTiles.x, W = settings.width;
		for (int y = my * tileSize.y; y < (my + 1) * tileSize.y; y++)
			std::memcpy(&pixels[size_t(W) * y + size_t(mx) * tileSize.x], &frame[3 * (size_t(W) * y + size_t(mx) * tileSize.x)], size_t(tileSize.x) * sizeof(vec3));
		finished = true;
	}

	// The whole frame without the tile protocol (every pixel, also the columns/rows the 40x10 tiling never reaches, Q28).
	void renderFrame() {
		std::atomic<bool> done{false};
		renderTile(camera, 0, done);
		std::memcpy(pixels, frame.data(), frame.size() * sizeof(float));
	}

	// OfflineEngine::coreLoop's scheduling (OfflineEngine.cpp:86-117): numberOfThreads caller-side threads over all tiles.
	void coreLoop() {
		const int tiles = numberOfTiles.x * numberOfTiles.y;
		std::atomic<int> next{0};
		const int nt = numberOfThreads < 16 ? numberOfThreads : 16;
		for (int t = 0; t < nt; t++)
			threadPool[t] = std::thread([&, t] {
				for (int i = next++; i < tiles; i = next++) {
					isThreadDone[t] = false;
					renderTile(camera, i, isThreadDone[t]);
				}
			});
		for (int t = 0; t < nt; t++) threadPool[t].join();
	}

	ne_b200_ctx* context() const { return ctx; }  // the first device's context
	int deviceCount() const { return ne_b200_multi_count(multi); }

private:
	ne_b200_multi* multi = nullptr;
	ne_b200_ctx* ctx = nullptr;
	std::vector<float> frame;
	bool frameReady = false;
	std::mutex frameMutex;
	static void check(int rc, const char* what) {
		if (rc != NE_B200_OK) throw std::runtime_error(std::string(what) + ": " + ne_b200_last_error());
	}
};

}  // namespace narval_b200
