// Wall-clock section timers used by the demos (reference kuiper/include/base/tick.h:6-16).
#ifndef KLLM_KUIPER_BASE_TICK_H_
#define KLLM_KUIPER_BASE_TICK_H_
#include <chrono>
#include <cstdio>
#include <iostream>

#define TICK(x) auto bench_##x = std::chrono::steady_clock::now();
#define TOCK(x)                                                                              \
  printf("%s: %lfs\n", #x,                                                                   \
         std::chrono::duration<double>(std::chrono::steady_clock::now() - bench_##x).count());
#endif  // KLLM_KUIPER_BASE_TICK_H_
