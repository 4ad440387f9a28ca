#include "eot_common.cuh"
using namespace eot;
extern "C" int eot_apply_bwd(const EotShape*, const float*, const float*, const float*, void*, size_t, float*, int, void*) {
  set_error("eot_apply_bwd: not built yet"); return EOT_ERR_BAD_SHAPE;
}
