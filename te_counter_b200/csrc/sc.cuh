// Single-cell path (placeholder until the kernels land).
#pragma once
#include "context.cuh"

struct ScState {};
inline void tec_ctx::free_sc() { delete sc; sc = nullptr; }

extern "C" int tec_sc_begin(tec_ctx* ctx, int, int, int64_t) { if (!ctx) return TEC_ERR_ARG; TEC_FAIL(TEC_ERR_UNIMPLEMENTED, "sc: not implemented yet"); }
extern "C" int tec_sc_push(tec_ctx* ctx, int64_t, const int32_t*, const int32_t*, const uint16_t*, const uint8_t*, const uint8_t*, const uint32_t*, const uint64_t*) { if (!ctx) return TEC_ERR_ARG; TEC_FAIL(TEC_ERR_UNIMPLEMENTED, "sc: not implemented yet"); }
extern "C" int tec_sc_push_dev(tec_ctx* ctx, int64_t, const int32_t*, const int32_t*, const uint16_t*, const uint8_t*, const uint8_t*, const uint32_t*, const uint64_t*) { if (!ctx) return TEC_ERR_ARG; TEC_FAIL(TEC_ERR_UNIMPLEMENTED, "sc: not implemented yet"); }
extern "C" int tec_sc_finalize(tec_ctx* ctx, int64_t, int64_t, int64_t, int64_t*, int64_t*) { if (!ctx) return TEC_ERR_ARG; TEC_FAIL(TEC_ERR_UNIMPLEMENTED, "sc: not implemented yet"); }
extern "C" int tec_sc_fetch(tec_ctx* ctx, int32_t*, uint32_t*, int64_t*, uint32_t*, int64_t*, int64_t*) { if (!ctx) return TEC_ERR_ARG; TEC_FAIL(TEC_ERR_UNIMPLEMENTED, "sc: not implemented yet"); }
extern "C" int tec_sc_select(tec_ctx* ctx, int64_t, uint32_t*, int64_t*) { if (!ctx) return TEC_ERR_ARG; TEC_FAIL(TEC_ERR_UNIMPLEMENTED, "sc: not implemented yet"); }
