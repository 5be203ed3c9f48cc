// placeholder while the primitives get their first GPU validation; replaced by the real engine
#include "../../include/sdxe.h"
#include "common.cuh"
using namespace sdxe;
extern "C" {
int sdxe_create(const sdxe_config*, sdxe_engine**) { set_last_error(__FILE__, __LINE__, "not implemented"); return -1; }
void sdxe_destroy(sdxe_engine*) {}
int sdxe_set_weight(sdxe_engine*, const char*, const void*, int, int, const int64_t*) { return -1; }
int64_t sdxe_param_count(const sdxe_engine*) { return 0; }
int sdxe_finalize(sdxe_engine*) { return -1; }
int sdxe_weight_blob(sdxe_engine*, void**, int64_t*) { return -1; }
int sdxe_unet_forward(sdxe_engine*, const void*, const void*, const void*, const void*, void*, int, int, int, int, int, void*) { return -1; }
int sdxe_vae_decode(sdxe_engine*, const void*, void*, int, int, int, int, void*) { return -1; }
}
