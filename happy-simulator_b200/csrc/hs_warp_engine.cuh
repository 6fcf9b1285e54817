#ifndef HS_WARP_ENGINE_CUH
#define HS_WARP_ENGINE_CUH
struct hs_engine;
static int hs_warp_launch(hs_engine *E, const hs_run_params *p, uint32_t ring, bool want_hash, bool want_rec);
#endif
