// logging stand-in: the reference logs through libimsux macros; the checker build discards them
// (set REF_OIP_LOG=1 to see them on stderr).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
inline void imsux_log_sink(const char *fmt, ...) {
    static const bool on = getenv("REF_OIP_LOG") != nullptr;
    if (!on) return;
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr);
}
enum { LSV_TRACE = 0 };
#define LOGT(...) imsux_log_sink(__VA_ARGS__)
#define LOGW(...) imsux_log_sink(__VA_ARGS__)
#define LOGE(...) imsux_log_sink(__VA_ARGS__)
#define LOGF(...) imsux_log_sink(__VA_ARGS__)
#define LOGX(level, raw) imsux_log_sink
