// Levelled stderr logging for the native core. The level comes from the LOG_LEVEL environment variable
// (TRACE | DEBUG | INFO | WARN | ERROR, default WARN) — the variable the reference's Rust core reads for its `tracing`
// subscriber (rust/bagua-core/bagua-core-py/src/lib.rs:542-547).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace bagua {
namespace log {
enum Level : int { TRACE = 0, DEBUG = 1, INFO = 2, WARN = 3, ERROR = 4 };

inline int parse_level(const char* v) {
    if (!v || !*v) return WARN;
    auto eq = [&](const char* s) { return strcasecmp(v, s) == 0; };
    if (eq("trace")) return TRACE;
    if (eq("debug")) return DEBUG;
    if (eq("info")) return INFO;
    if (eq("warn") || eq("warning")) return WARN;
    if (eq("error")) return ERROR;
    return WARN;
}
inline int& threshold() {
    static int lvl = parse_level(getenv("LOG_LEVEL"));
    return lvl;
}
inline bool enabled(int lvl) { return lvl >= threshold(); }
inline const char* name(int lvl) {
    static const char* n[] = {"TRACE", "DEBUG", "INFO", "WARN", "ERROR"};
    return n[lvl < 0 ? 0 : (lvl > 4 ? 4 : lvl)];
}
}  // namespace log
}  // namespace bagua

#define BAGUA_LOG(lvl, ...)                                                                                             \
    do {                                                                                                                \
        if (::bagua::log::enabled(::bagua::log::lvl)) {                                                                 \
            const double _t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); \
            std::fprintf(stderr, "[bagua %s %.6f] ", ::bagua::log::name(::bagua::log::lvl), _t);                        \
            std::fprintf(stderr, __VA_ARGS__);                                                                          \
            std::fputc('\n', stderr);                                                                                   \
        }                                                                                                               \
    } while (0)
