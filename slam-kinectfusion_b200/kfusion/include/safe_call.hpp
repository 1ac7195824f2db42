// safe_call.hpp -- error convention of the host facade (mirrors kfusion/include/safe_call.hpp).
// The reference prints "cuda error: ..." and continues; the C-ABI returns codes, and the facade
// keeps the print-and-continue behaviour through kfbSafeCall so call sites read the same.
#pragma once
#include <iostream>
#include "../../../include/kfb200.h"
#define kfbSafeCall(ctx, expr) kf::___kfbSafeCall((ctx), (expr), __FILE__, __LINE__)
namespace kf
{
static inline int ___kfbSafeCall(kfb_ctx *ctx, int rc, const char *file, const int line)
{
    if (rc != KFB_OK) std::cout << "cuda error: " << kfb_last_error_string(ctx) << "\t" << file << ":" << line << std::endl;
    return rc;
}
} // namespace kf
