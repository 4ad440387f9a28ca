#include "eot_common.cuh"
using namespace eot;
extern "C" int score_workspace_bytes(const ScoreShape*, size_t*) { set_error("not built yet"); return EOT_ERR_BAD_SHAPE; }
extern "C" int score_max_fwd(const ScoreShape*, const float* const*, const float* const*, const float*, float*, int32_t*, int32_t*, void*, size_t, void*) { set_error("not built yet"); return EOT_ERR_BAD_SHAPE; }
extern "C" int score_max_bwd(const ScoreShape*, const float* const*, const float*, const float*, float* const*, float*, float*, void*, size_t, void*) { set_error("not built yet"); return EOT_ERR_BAD_SHAPE; }
