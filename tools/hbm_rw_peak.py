"""Pure-write / pure-read / copy HBM bandwidth on this B200 (the Gram kernel is write-dominated: 4 B/pair)."""
import torch
dev = torch.device("cuda:0")
n = 6553600000 // 4  # 6.55 GB of fp32, the size of one 16384 x 100000 Gram block
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print("fill_ (st.global)   : %.3f ms  %.0f GB/s write" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: x.zero_());    print("zero_ (memset)      : %.3f ms  %.0f GB/s write" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x));   print("copy_ (read+write)  : %.3f ms  %.0f GB/s total" % (ms, 2 * n * 4 / ms / 1e6))
ms = t(lambda: x.sum());      print("sum   (read)        : %.3f ms  %.0f GB/s read" % (ms, n * 4 / ms / 1e6))
