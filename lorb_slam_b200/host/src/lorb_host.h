// lorb_host.h — glue shared by the drop-in Matcher / BA / ORBextractor bodies: one lorb_ctx
// per calling thread (borrowed from a process-wide pool) (the reference runs Matcher and the pose-only BA on the
// tracking thread and local BA on the mapper thread, example/main.cpp:36-37),
// and a fatal-error policy: the reference has no error channel,
// and a silent "0 matches" would hide a broken GPU path, so any non-zero C-ABI
// status aborts with the library's message.  There is no CPU fallback.
#ifndef LORB_HOST_H
#define LORB_HOST_H
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../../include/lorb_cuda.h"

namespace lorb_host {

inline void die(const char* what, int rc) {
  std::fprintf(stderr, "[lorb] %s failed (status %d): %s\n", what, rc, lorb_last_error());
  std::abort();
}

// Contexts are pooled, not owned by threads: the reference starts fresh std::threads for every
// frame (Frame::Frame, src/frame.cpp:42-45: one per image), and a context carries warm state worth
// keeping -- stream, grow-only device / pinned buffers, the captured CUDA graph of the extractor.
// A thread borrows a context on first use and hands it back when it exits; the pool destroys them
// at process exit.  One context is never used by two threads at a time.
class CtxPool {
 public:
  lorb_ctx* acquire() {
    {
      std::lock_guard<std::mutex> g(m_);
      if (!free_.empty()) {
        lorb_ctx* c = free_.back();
        free_.pop_back();
        return c;
      }
    }
    lorb_ctx* c = nullptr;
    const char* dev = std::getenv("LORB_DEVICE");
    const int rc = lorb_ctx_create(dev ? std::atoi(dev) : 0, &c);
    if (rc != LORB_OK) die("lorb_ctx_create", rc);
    std::lock_guard<std::mutex> g(m_);
    ++created_;
    return c;
  }
  int created() {
    std::lock_guard<std::mutex> g(m_);
    return created_;
  }
  void release(lorb_ctx* c) {
    std::lock_guard<std::mutex> g(m_);
    free_.push_back(c);
  }
  // Explicit teardown for hosts that want the device memory back before exit (no thread may be
  // inside a Matcher / BA / ORBextractor call).  There is deliberately NO destructor doing this: the
  // pool object is never destroyed (see pool()), because CUDA's own atexit teardown may already have
  // run when static destructors fire, and a thread that exits after main must still find a live pool.
  void shutdown() {
    std::lock_guard<std::mutex> g(m_);
    for (lorb_ctx* c : free_) lorb_ctx_destroy(c);
    free_.clear();
  }

 private:
  std::mutex m_;
  std::vector<lorb_ctx*> free_;
  int created_ = 0;
};

inline CtxPool& pool() {
  // one pool per process (inline function: shared by every translation unit), leaked on purpose:
  // contexts left in it at exit are reclaimed by the driver with the process
  static CtxPool* p = new CtxPool();
  return *p;
}

inline void shutdown() { pool().shutdown(); }

struct ThreadCtx {
  lorb_ctx* ctx = nullptr;
  ~ThreadCtx() {
    if (ctx) pool().release(ctx);
  }
};

inline lorb_ctx* ctx() {
  static thread_local ThreadCtx t;
  if (!t.ctx) t.ctx = pool().acquire();
  return t.ctx;
}

#define LORB_HOST_CALL(expr)                       \
  do {                                             \
    const int _rc = (expr);                        \
    if (_rc != LORB_OK) lorb_host::die(#expr, _rc); \
  } while (0)

}  // namespace lorb_host
#endif
