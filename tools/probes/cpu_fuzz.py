"""CPU probe (not a test): mutated and truncated files through the host build of the kernel code (tests/hostemu), optionally
under AddressSanitizer + UBSan.  The section decoders, palette / pass / upsampling code are the same headers the CUDA kernels
compile, so an out-of-bounds access found here is one on the device too.

  python tools/probes/cpu_fuzz.py [--asan] [N per file] [files ...]      (default: the palette / LZ77 / progressive / upsampled
                                                                          files of tests/_cache, 120 mutations each)
--asan builds tests/hostemu with -fsanitize=address,undefined into /tmp and re-runs itself with libasan preloaded."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(path, n):
    import numpy as np
    import hostemu_lib as H
    if os.environ.get("JXLB_FUZZ_LIB"):
        H.build = lambda: os.environ["JXLB_FUZZ_LIB"]
    base = open(path, "rb").read()
    rng = np.random.default_rng(7)
    ok = err = 0
    for k in range(n):
        d = bytearray(base)
        m = rng.integers(0, 4)
        if m == 0:
            for _ in range(int(rng.integers(1, 4))):
                d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        elif m == 1:
            d[int(rng.integers(0, min(len(d), 400)))] = int(rng.integers(0, 256))
        elif m == 2:
            d = d[: int(rng.integers(1, len(d)))]
        else:
            i = int(rng.integers(0, len(d) - 8))
            d[i:i + 8] = bytes(rng.integers(0, 256, 8).astype(np.uint8))
        try:
            e = H.Decoded(bytes(d))
        except RuntimeError:
            err += 1
            continue
        try:
            if e.status == 0:
                e.render()
                ok += 1
            else:
                err += 1
        except RuntimeError:
            err += 1
        e.close()
    print("%s: %d decoded, %d refused / failed cleanly" % (os.path.basename(path), ok, err), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(args[1], int(args[2]))
    env = dict(os.environ)
    if args and args[0] == "--asan":
        args = args[1:]
        import hostemu_lib as H
        out = "/tmp/libhostemu_asan.so"
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fPIC", "-shared",
                               "-Wno-unknown-pragmas", "-o", out] + H.SRC, cwd=ROOT)
        env.update(JXLB_FUZZ_LIB=out, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1",
                   LD_PRELOAD=subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip())
    n = 120
    if args and args[0].isdigit():
        n = int(args[0])
        args = args[1:]
    files = args or sorted(sum((glob.glob(os.path.join(ROOT, "tests", "_cache", p)) for p in ("pal_*.jxl", "e1_*.jxl", "prog_*_d1_e7*.jxl", "up2_*_d12_*.jxl")), []))
    bad = 0
    for f in files:
        r = subprocess.run([sys.executable, __file__, "--child", f, str(n)], env=env, capture_output=True, text=True, timeout=1800)
        print(r.stdout.strip() or "(no output)")
        if r.returncode:
            bad += 1
            print("  CRASH rc=%d\n%s" % (r.returncode, r.stderr[-1500:]))
    print("files %d, crashed %d" % (len(files), bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
