// safe_call.hpp -- error convention of the host facade.  The reference's wrapper (kfusion/include/safe_call.hpp:8-14)
// prints "cuda error: <text>\t<file>:<line>" to stdout and carries on; the C-ABI returns codes instead, and the
// facade keeps the print-and-continue behaviour through kfbSafeCall so that a failing device call is as visible
// (and as non-fatal) as it is in the reference.  The code is handed back for the callers that do care.
//
// The frame loop itself (upload, front end, ICP, integrate, raycast, composite) uses kfbCheck instead: a device or
// transport error there is NOT a tracking failure -- carrying on would reset or silently corrupt the map -- so it is
// raised as kf::DeviceError and the frame is abandoned with the volume and the pose history untouched.
#pragma once
#include <iostream>
#include <stdexcept>
#include <string>
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
inline void check(kfb_ctx *ctx, int code, const char *where, int line);
} // namespace detail
struct DeviceError : std::runtime_error
{
    int code;
    DeviceError(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};
inline void detail::check(kfb_ctx *ctx, int code, const char *where, int line)
{
    if (code == KFB_OK) return;
    const char *what = kfb_last_error_string(ctx);
    throw DeviceError(code, std::string("kfb200 error ") + std::to_string(code) + ": " + (what ? what : "?") + " @ " + where + ":" + std::to_string(line));
}
} // namespace kf
#define kfbSafeCall(ctx, call) kf::detail::report((ctx), (call), __FILE__, __LINE__)
#define kfbCheck(ctx, call) kf::detail::check((ctx), (call), __FILE__, __LINE__)
