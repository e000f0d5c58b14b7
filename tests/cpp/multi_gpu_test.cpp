// The single-process multi-GPU path of the C ABI (ne_b200_create_multi, csrc/ne_multi.cu) driven the way NarvalEngine's
// one-process editor would (SceneEditor::startOffEngine, src/SceneEditor.cpp:590-604): B200OfflineEngine with a device
// list renders the frame split by sample index over the devices; the result must be the one-device frame up to the
// order of fp32 additions. usage: multi_gpu_test scene.json resources_dir dev0,dev1[,...]   (a device may repeat)
// Prints "OK ..." and exits 0; exits 3 with the library's message when no CUDA device is present.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ne_b200_offline_engine.hpp"

int main(int argc, char** argv) {
	if (argc < 4) { fprintf(stderr, "usage: %s scene.json resources_dir dev0,dev1,...\n", argv[0]); return 2; }
	std::vector<int> devices;
	for (const char* p = argv[3]; *p;) {
		devices.push_back(int(strtol(p, const_cast<char**>(&p), 10)));
		if (*p == ',') p++;
	}
	ne_b200_scene_file* file = nullptr;
	if (ne_b200_scene_file_load(argv[1], argv[2], &file) != NE_B200_OK) { fprintf(stderr, "load: %s\n", ne_b200_last_error()); return 2; }
	ne_b200_camera cam;
	ne_b200_render_settings st;
	ne_b200_scene_file_camera(file, &cam);
	ne_b200_scene_file_settings(file, &st);
	int rc = 0;
	try {
		narval_b200::B200OfflineEngine one(cam, st, ne_b200_scene_file_desc(file), devices[0]);
		one.renderFrame();
		narval_b200::B200OfflineEngine many(cam, st, ne_b200_scene_file_desc(file), devices);
		if (many.deviceCount() != int(devices.size())) { fprintf(stderr, "device count\n"); return 1; }
		many.coreLoop();  // SceneEditor-style tile protocol on top of the multi-GPU frame
		many.renderFrame();
		const size_t n = size_t(st.width) * st.height * 3;
		double sum = 0, maxRel = 0, mean = 0;
		for (size_t i = 0; i < n; i++) mean += one.linear[i];
		mean /= double(n);
		for (size_t i = 0; i < n; i++) {
			double d = std::fabs(double(one.linear[i]) - double(many.linear[i]));
			double rel = d / (std::fabs(double(one.linear[i])) + 1e-3 * mean);
			if (rel > maxRel) maxRel = rel;
			sum += many.linear[i];
		}
		if (!(mean > 0) || maxRel > 2e-3) { fprintf(stderr, "multi-GPU frame differs from the one-GPU frame: max rel %g (mean %g)\n", maxRel, mean); rc = 1; }
		else printf("OK %d devices %dx%d spp %d mean %.6f max_rel %.3g\n", many.deviceCount(), st.width, st.height, st.spp, sum / double(n), maxRel);
	} catch (const std::exception& e) {
		fprintf(stderr, "%s\n", e.what());
		rc = 3;
	}
	ne_b200_scene_file_free(file);
	return rc;
}
