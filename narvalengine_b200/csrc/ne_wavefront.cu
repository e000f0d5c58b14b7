// ne_wavefront.cu — production wavefront renderer (placeholder until the queues land).
#include "ne_ctx.h"
namespace ne {
int wavefront_render(ne_b200_ctx*, int, int, int, uint64_t, uint32_t) {
	set_error("wavefront renderer not built yet: pass NE_B200_RENDER_MEGAKERNEL");
	return NE_B200_ERR_UNSUPPORTED;
}
void wavefront_free(ne_b200_ctx*) {}
}  // namespace ne
