// emu_cuda.cpp -- fiber scheduler of the host warp emulator (TEST INFRASTRUCTURE, see emu_cuda.h).
#include "emu_cuda.h"

#include <stdarg.h>

namespace emu {

thread_local Warp* g_warp = nullptr;

// Minimal x86-64 context switch: push callee-saved registers, swap stack pointers, pop.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

int lane() { return g_warp->cur; }

static void to_main() {
  Warp* w = g_warp;
  const int me = w->cur;
  emu_switch(&w->lanes[me].sp, w->main_sp);
}

void yield() { to_main(); }

[[noreturn]] void fail(const char* fmt, ...) {
  Warp* w = g_warp;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(w->fail_msg, sizeof(w->fail_msg), fmt, ap);
  va_end(ap);
  w->failed = true;
  to_main();
  abort();   // never resumed
}

static void lane_entry() {
  Warp* w = g_warp;
  const int me = w->cur;
  w->body(me);
  w->lanes[me].done = true;
  to_main();
  abort();
}

// All live lanes meet here.  The scheduler (run_warp) resumes lanes round-robin; a lane that arrives
// first simply yields until the generation changes.
void rendezvous(int opcode, const void* site) {
  Warp* w = g_warp;
  const int me = w->cur;
  w->op[me] = opcode;
  w->site[me] = site;
  const uint64_t my_gen = w->gen;
  w->arrived++;
  w->n_sync++;
  if (w->arrived == kLanes) {
    for (int i = 1; i < kLanes; ++i)
      if (w->op[i] != w->op[0]) fail("divergent warp intrinsic: lane 0 op %d, lane %d op %d", w->op[0], i, w->op[i]);
    w->arrived = 0;
    w->gen++;
    return;
  }
  while (w->gen == my_gen) to_main();
}

void check_global(const void* p, size_t n, bool write) {
  Warp* w = g_warp;
  if (w->n_regions == 0) return;
  const uint8_t* a = (const uint8_t*)p;
  for (int i = 0; i < w->n_regions; ++i) {
    const Region& r = w->regions[i];
    // vector accesses may touch the 16-byte granules that contain valid bytes
    const uint8_t* lo = (const uint8_t*)((uintptr_t)r.lo & ~(uintptr_t)15);
    const uint8_t* hi = (const uint8_t*)(((uintptr_t)r.hi + 15) & ~(uintptr_t)15);
    if (a >= lo && a + n <= hi) {
      if (write && (!r.writable || a < r.lo || a + n > r.hi)) fail("global write [%p,+%zu) outside writable region", p, n);
      return;
    }
  }
  fail("global access [%p,+%zu) outside every registered region", p, n);
}

void run_warp(Warp& w, size_t smem_bytes, std::function<void(int)> body) {
  g_warp = &w;
  w.body = std::move(body);
  w.smem_bytes = smem_bytes;
  w.smem = (uint8_t*)aligned_alloc(128, (smem_bytes + 127) & ~(size_t)127);
  memset(w.smem, 0xcd, smem_bytes);
  w.failed = false;
  w.arrived = 0;
  w.gen = 0;
  for (int i = 0; i < kLanes; ++i) {
    Fiber& f = w.lanes[i];
    f.done = false;
    f.stack = (uint8_t*)aligned_alloc(64, kStackBytes);
    // initial frame: six callee-saved registers + return address (lane_entry); keep the ABI's
    // 16-byte alignment at function entry (rsp % 16 == 8 after the return address is popped)
    uint64_t* top = (uint64_t*)(f.stack + kStackBytes - 64);
    top = (uint64_t*)((uintptr_t)top & ~(uintptr_t)15);
    *(--top) = 0;                              // fake return address of lane_entry
    *(--top) = (uint64_t)(uintptr_t)&lane_entry;
    for (int k = 0; k < 6; ++k) *(--top) = 0;
    f.sp = top;
  }
  int live = kLanes;
  while (live > 0 && !w.failed) {
    int progressed = 0;
    for (int i = 0; i < kLanes && !w.failed; ++i) {
      if (w.lanes[i].done) continue;
      w.cur = i;
      emu_switch(&w.main_sp, w.lanes[i].sp);
      ++progressed;
      if (w.lanes[i].done) {
        --live;
        if (w.arrived != 0 && live > 0) {
          // a lane left while others wait inside a full-mask intrinsic
          snprintf(w.fail_msg, sizeof(w.fail_msg), "lane %d exited while %d lanes wait in a warp intrinsic", i, w.arrived);
          w.failed = true;
        }
      }
    }
    if (!progressed) break;
  }
  for (int i = 0; i < kLanes; ++i) free(w.lanes[i].stack);
  free(w.smem);
  w.smem = nullptr;
  g_warp = nullptr;
}

}  // namespace emu
