// lorb_host.h — glue shared by the drop-in Matcher / BA bodies: one lorb_ctx
// per calling thread (the reference runs Matcher and the pose-only BA on the
// tracking thread and local BA on the mapper thread, example/main.cpp:36-37),
// and a fatal-error policy: the reference has no error channel,
// and a silent "0 matches" would hide a broken GPU path, so any non-zero C-ABI
// status aborts with the library's message.  There is no CPU fallback.
#ifndef LORB_HOST_H
#define LORB_HOST_H
#include <cstdio>
#include <cstdlib>

#include "../../../include/lorb_cuda.h"

namespace lorb_host {

struct ThreadCtx {
  lorb_ctx* ctx = nullptr;
  ~ThreadCtx() {
    if (ctx) lorb_ctx_destroy(ctx);
  }
};

inline void die(const char* what, int rc) {
  std::fprintf(stderr, "[lorb] %s failed (status %d): %s\n", what, rc, lorb_last_error());
  std::abort();
}

inline lorb_ctx* ctx() {
  static thread_local ThreadCtx t;
  if (!t.ctx) {
    const char* dev = std::getenv("LORB_DEVICE");
    const int rc = lorb_ctx_create(dev ? std::atoi(dev) : 0, &t.ctx);
    if (rc != LORB_OK) die("lorb_ctx_create", rc);
  }
  return t.ctx;
}

#define LORB_HOST_CALL(expr)                       \
  do {                                             \
    const int _rc = (expr);                        \
    if (_rc != LORB_OK) lorb_host::die(#expr, _rc); \
  } while (0)

}  // namespace lorb_host
#endif
