#!/usr/bin/env python
"""Host link probe: pinned H2D / D2H copy rates alone and together (what bounds every end-to-end number of bench.py)."""
import time

import torch


def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device='cuda')
    d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=5, piece=n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for lo in range(0, n, piece):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in[lo:lo + piece].copy_(h_in[lo:lo + piece], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[lo:lo + piece].copy_(d_out[lo:lo + piece], non_blocking=True)
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t0) / 1e9

    run(True, True, 1)
    print('H2D alone      %.1f GB/s' % run(True, False))
    print('D2H alone      %.1f GB/s' % run(False, True))
    print('both at once   %.1f GB/s each' % run(True, True))
    print('both, 64 MiB pieces   %.1f GB/s each' % run(True, True, piece=64 << 20))
    print('D2H, 4 MiB pieces     %.1f GB/s' % run(False, True, piece=4 << 20))


if __name__ == '__main__':
    main()
