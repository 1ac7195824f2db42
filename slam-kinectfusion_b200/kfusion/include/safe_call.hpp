// safe_call.hpp -- error convention of the host facade.  The reference's wrapper (kfusion/include/safe_call.hpp:8-14)
// prints "cuda error: <text>\t<file>:<line>" to stdout and carries on; the C-ABI returns codes instead, and the
// facade keeps the print-and-continue behaviour through kfbSafeCall so that a failing device call is as visible
// (and as non-fatal) as it is in the reference.  The code is handed back for the callers that do care.
#pragma once
#include <iostream>
#include "../../../include/kfb200.h"

namespace kf
{
namespace detail
{
inline int report(kfb_ctx *ctx, int code, const char *where, int line)
{
    if (code == KFB_OK) return code;
    const char *what = kfb_last_error_string(ctx);
    std::cout << "cuda error: " << (what ? what : "?") << "\t" << where << ":" << line << std::endl;
    return code;
}
} // namespace detail
} // namespace kf
#define kfbSafeCall(ctx, call) kf::detail::report((ctx), (call), __FILE__, __LINE__)
