"""Stall monitor for multi-GPU runs (diagnostic tooling, not on the product path).

    import tools.hangwatch as hw;  hw.start(stall_s=25);  hw.beat("step 3")  # call beat() whenever progress is made

When no beat() arrives for `stall_s` seconds the monitor thread writes a report to stderr and to
gpurun_out/hang_rank<R>.log and ends the process with exit code 7, so that a wedged collective or kernel shows WHERE
it is wedged instead of sitting in the NCCL watchdog for minutes:

  * the Python stack of every thread (faulthandler) - a host thread blocked in a CUDA call shows the calling line;
  * the op trace kept by videonavqa_b200.ops (ops.TRACE): one CUDA event per C-ABI launch, queried here to name the
    first launch that has not completed on the device;
  * `cuda-gdb -batch` attached to this process: the kernels resident on the GPU and the native stacks
    (skipped with HANGWATCH_GDB=0).
"""
import faulthandler
import os
import subprocess
import sys
import threading
import time

_last = [time.time(), "start"]
_started = False
_dbg = [None]        # pinned int64 tensor handed to vnqa_debug_install when the -DVNQA_DEBUG_MBAR library is loaded
REC, SLOTS = 32, 148 * 8
ROLE = {1: "producer waits empty", 2: "mma waits acc_empty", 3: "mma waits full", 4: "epilogue waits acc_full"}


def _install_debug_buffer():
    """The bring-up build of the library (VNQA_LIB_PATH=build/libvnqa_b200_dbg.so) records stalled mbarrier waits into
    pinned host memory; nothing happens with the release library."""
    try:
        import ctypes
        import torch
        from videonavqa_b200 import _lib
        if _lib.lib().vnqa_debug_build() != 1:
            return
        buf = torch.zeros(SLOTS * REC, dtype=torch.int64).pin_memory()
        _lib.lib().vnqa_debug_install(ctypes.c_void_p(buf.data_ptr()))
        _dbg[0] = buf
        print(f"hangwatch: mbarrier debug buffer installed ({buf.numel() * 8} bytes pinned)", file=sys.stderr, flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"hangwatch: no debug buffer ({e!r})", file=sys.stderr, flush=True)


def _dump_debug_records(say):
    buf = _dbg[0]
    if buf is None:
        return
    import numpy as np
    a = buf.numpy().view(np.uint64).reshape(SLOTS, REC)
    rows = [r for r in a if int(r[0]) == 0x564E5141]
    say(f"--- mbarrier debug records: {len(rows)} stalled waits")
    for r in sorted(rows, key=lambda r: (int(r[1]) >> 32, int(r[1]) & 0xFFFFFFFF)):
        blk, tid = int(r[1]) >> 32, int(r[1]) & 0xFFFFFFFF
        tag, par, bidx = int(r[2]) >> 32, (int(r[2]) >> 16) & 0xFFFF, int(r[2]) & 0xFFFF
        info = int(r[3])
        smid, crank, rounds = int(r[4]) >> 32, (int(r[4]) >> 16) & 0xFFFF, int(r[4]) & 0xFFFF
        bars = " ".join(f"{int(x):016x}" for x in r[6:6 + 22])
        say(f"  block {blk:3d} rank {crank} sm {smid:3d} warp {tid >> 5} [{ROLE.get(tag, tag)}] bar#{bidx} parity {par} "
            f"item {info >> 20} sub {info & 0xFFFFF} rounds {rounds}\n      bars: {bars}")


def beat(what=""):
    _last[0] = time.time()
    _last[1] = what


def _report(stall_s, out_dir):
    rank = os.environ.get("RANK", "0")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "hang_" + os.environ.get("HANGWATCH_TAG", f"rank{rank}") + ".log")
    with open(path, "w") as f:
        def say(msg):
            for fh in (f, sys.stderr):
                print(msg, file=fh, flush=True)
        say(f"=== hangwatch rank {rank} pid {os.getpid()}: no progress for {stall_s:.0f} s after '{_last[1]}' "
            f"at {time.strftime('%H:%M:%S')}")
        _dump_debug_records(say)
        say("--- python stacks")
        faulthandler.dump_traceback(file=f, all_threads=True)
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        try:
            from videonavqa_b200 import ops
            tr = list(ops.TRACE or [])
            say(f"--- op trace: {len(tr)} launches recorded")
            first_pending = None
            for i, (what, ev, stream) in enumerate(tr):
                if not ev.query():
                    first_pending = i
                    break
            if first_pending is None:
                say("every traced launch has completed on the device")
            else:
                lo = max(0, first_pending - 6)
                for i in range(lo, min(len(tr), first_pending + 8)):
                    what, ev, stream = tr[i]
                    say(f"  [{i}] {what:28s} stream {stream:#x} {'done' if ev.query() else 'PENDING'}")
        except Exception as e:  # noqa: BLE001
            say(f"op trace unavailable: {e!r}")
        if os.environ.get("HANGWATCH_GDB", "1") != "0":
            say("--- cuda-gdb")
            try:
                r = subprocess.run(["cuda-gdb", "-p", str(os.getpid()), "-batch", "-ex", "info cuda kernels", "-ex",
                                    "info cuda blocks", "-ex", "info cuda threads", "-ex", "info cuda warps", "-ex",
                                    "info threads"], capture_output=True, text=True,
                                   timeout=90)
                say(r.stdout[-12000:])
                say(r.stderr[-3000:])
            except Exception as e:  # noqa: BLE001
                say(f"cuda-gdb failed: {e!r}")
    time.sleep(float(os.environ.get("HANGWATCH_LINGER_S", "40")))   # let the other ranks finish their reports
    os._exit(7)


def start(stall_s=30.0, out_dir=None):
    global _started
    if _started:
        return
    _started = True
    out_dir = out_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    beat("hangwatch armed")
    _install_debug_buffer()

    def loop():
        while True:
            time.sleep(1.0)
            if time.time() - _last[0] > stall_s:
                _report(time.time() - _last[0], out_dir)
    threading.Thread(target=loop, daemon=True, name="hangwatch").start()
