// TICK(name) ... TOCK(name): print the wall-clock seconds a section took (the demos time their
// generation loop with these).
#ifndef KLLM_KUIPER_BASE_TICK_H_
#define KLLM_KUIPER_BASE_TICK_H_
#include <chrono>
#include <cstdio>

namespace base {
struct SectionTimer {
  std::chrono::steady_clock::time_point begin = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count(); }
};
}  // namespace base
#define TICK(x) const base::SectionTimer bench_##x;
#define TOCK(x) std::printf("%s: %lfs\n", #x, bench_##x.seconds());
#endif  // KLLM_KUIPER_BASE_TICK_H_
