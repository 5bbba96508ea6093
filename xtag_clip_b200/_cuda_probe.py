"""Start-up guard for GPU entry points (tests, smoke, bench): on a freshly provisioned box the first cuInit of a
process can fail for a few seconds ("CUDA driver initialization failed") and the failure is sticky inside that
process.  `wait_for_cuda` therefore probes in SUBPROCESSES until one sees a device (or the deadline passes) before the
caller touches torch.cuda itself.  On a machine without NVIDIA device nodes it returns False immediately."""
from __future__ import annotations

import os
import subprocess
import sys
import time


def has_device_nodes() -> bool:
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def wait_for_cuda(timeout_s: float = 90.0) -> bool:
    if not has_device_nodes():
        return False
    t0 = time.time()
    probe = "import sys, torch; sys.exit(0 if torch.cuda.is_available() and torch.cuda.device_count() > 0 else 3)"
    while True:
        try:
            if subprocess.run([sys.executable, "-c", probe], timeout=180, stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL).returncode == 0:
                return True
        except Exception:
            pass
        if time.time() - t0 > timeout_s:
            return False
        time.sleep(3.0)
