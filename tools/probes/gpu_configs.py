"""GPU probe (not a test): the BASELINE configs other than the bench workload, timed through the public API with host
buffers, the reference timed beside them on the box's CPU.  Prints one line per config."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jxl_coder_b200 as J  # noqa: E402
from oracle import gen_inputs, refjxl  # noqa: E402


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.time()
        fn()
        ts.append(time.time() - t)
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    # configs[0]: 512x512 lossless RGBA8
    d = gen_inputs.c1_image()
    b, m = best(lambda: J.JxlCoder.decode(d, 2))
    rb, rm = best(lambda: refjxl.decode_sampled(d, cfg=2), 3)
    print("C1 512x512 lossless -> RGBA8: gpu best %.2f ms median %.2f ms | reference best %.2f ms" % (b * 1e3, m * 1e3, rb * 1e3), flush=True)
    # configs[3]: 8K -> 1080p 1010102
    d = gen_inputs.c4_image()
    b, m = best(lambda: J.JxlCoder.decode_sampled(d, 1920, 1080, 5, 1, 4))
    print("C4 timings", J.last_batch_timings())
    rb, rm = best(lambda: refjxl.decode_sampled(d, w=1920, h=1080, cfg=5, scale_mode=1, filt=4), 2)
    mp = 7680 * 4320 / 1e6
    print("C4 8K -> 1080p 1010102: gpu best %.1f ms (%.0f MP/s) median %.1f ms | reference best %.1f ms (%.0f MP/s)" %
          (b * 1e3, mp / b, m * 1e3, rb * 1e3, mp / rb), flush=True)
    b, m = best(lambda: J.JxlCoder.decode(d, 2))
    print("C4 8K decode only -> RGBA8: gpu best %.1f ms (%.0f MP/s)" % (b * 1e3, mp / b), flush=True)
    # configs[2]: 256 x 1080p -> F16
    ds = [gen_inputs.c3_image(i % 4) for i in range(256)]
    def c3():
        for bmp in J.decode_batch(ds, config=3, keep_native=True):
            bmp.free()
    b, m = best(c3, 3)
    mp = 256 * 1920 * 1080 / 1e6
    print("C3 timings", J.last_batch_timings())
    t = time.time()
    for i in range(8):
        refjxl.decode_sampled(ds[i], cfg=3)
    rt = (time.time() - t) / 8
    print("C3 256 x 1080p -> F16 (host out): gpu best %.1f ms (%.0f MP/s) median %.1f ms | reference %.1f ms per image serial (%.0f MP/s)" %
          (b * 1e3, mp / b, m * 1e3, rt * 1e3, 1920 * 1080 / 1e6 / rt), flush=True)


def c5():
    # configs[4]: every frame of a 120-frame 1024x1024 RGBA lossy animation through JxlAnimatedImage.getFrame(i)
    d = gen_inputs.c5_animation()
    a = J.JxlAnimatedImage(d, 2)
    n = a.number_of_frames
    a.get_frame(0)
    t = time.time()
    for i in range(n):
        a.get_frame(i)
    dt = time.time() - t
    a.close()
    ra = refjxl.Anim(d, cfg=2)
    rt = {}
    for i in (0, 10, 40, n - 1):
        t0 = time.time()
        ra.frame(i)
        rt[i] = time.time() - t0
    ra.close()
    print("C5 %d-frame 1024x1024 RGBA lossy animation, getFrame(i) for every i: gpu %.1f ms per frame (%.0f MP/s) | reference getFrame(i) "
          "re-decodes frames 0..i: %s ms" % (n, dt / n * 1e3, n * 1.048576 / dt, {k: round(v * 1e3, 1) for k, v in rt.items()}), flush=True)


if __name__ == "__main__":
    c5()
    main()
