"""Append-only checkpoint store of a run.

The reference appends every step to emcee's HDF5 backend (model.py:204-220) and resumes from it
(model.py:235-253).  Here a run is a directory of blocks -- one `.npz` per checkpoint interval, written
once, never rewritten -- plus a small `state.json`; total I/O is O(steps), every file appears atomically
(temporary name + os.replace), and the writing happens on a background thread while the GPU samples the
next block.  A legacy single-file checkpoint (`chain_state.npz`) is read as the first block.
"""
import json
import os
import queue
import threading
from pathlib import Path

import numpy as np


class ChainStore:
    def __init__(self, root, legacy=None):
        self.root = Path(root)
        self.legacy = Path(legacy) if legacy is not None else None
        self._q = None
        self._thread = None
        self._error = None

    # ------------------------------------------------------------------ reading
    def _state_path(self):
        return self.root / "state.json"

    def exists(self):
        return self._state_path().is_file() or (self.legacy is not None and self.legacy.is_file())

    def reset(self):
        self.close()
        if self.root.is_dir():
            for p in self.root.iterdir():
                p.unlink()
            self.root.rmdir()
        if self.legacy is not None and self.legacy.is_file():
            self.legacy.unlink()

    def load(self):
        """-> dict(iteration, converged, chain (W, it, ndim) or None, log_prob (it, W) or None)."""
        chains, logps, it, conv = [], [], 0, False
        if self.legacy is not None and self.legacy.is_file():
            with np.load(self.legacy) as z:
                chains.append(z["chain"])
                logps.append(z["log_prob"])
                it = int(z["iteration"])
        if self._state_path().is_file():
            st = json.loads(self._state_path().read_text())
            for k in range(int(st["blocks"])):
                with np.load(self.root / f"block_{k:06d}.npz") as z:
                    if int(z["first"]) != it:
                        raise ValueError(f"checkpoint block {k} starts at iteration {int(z['first'])}, expected {it}")
                    chains.append(z["chain"])
                    logps.append(z["log_prob"])
                    it += z["chain"].shape[1]
            if it != int(st["iteration"]):
                raise ValueError("checkpoint state and blocks disagree")
            conv = bool(st.get("converged", False))
        if not chains:
            return {"iteration": 0, "converged": False, "chain": None, "log_prob": None}
        return {"iteration": it, "converged": conv, "chain": np.concatenate(chains, axis=1),
                "log_prob": np.concatenate(logps, axis=0)}

    # ------------------------------------------------------------------ writing
    def _nblocks(self):
        if self._state_path().is_file():
            return int(json.loads(self._state_path().read_text())["blocks"])
        return 0

    def _write(self, k, first, chain, log_prob, iteration, converged):
        self.root.mkdir(parents=True, exist_ok=True)
        tmp = self.root / f".block_{k:06d}.tmp.npz"
        np.savez(tmp, first=first, chain=chain, log_prob=log_prob)
        os.replace(tmp, self.root / f"block_{k:06d}.npz")
        tmp = self.root / ".state.tmp"
        tmp.write_text(json.dumps({"blocks": k + 1, "iteration": int(iteration), "converged": bool(converged)}))
        os.replace(tmp, self._state_path())

    def _worker(self):
        while True:
            job = self._q.get()
            if job is None:
                return
            try:
                self._write(*job)
            except Exception as ex:  # surfaced by the next append() / close()
                self._error = ex

    def append(self, first, chain, log_prob, iteration, converged=False):
        """Queue one block (chain (W, n, ndim), log_prob (n, W), covering iterations [first, iteration))."""
        if self._error is not None:
            raise self._error
        if self._thread is None:
            self._next = self._nblocks()
            self._q = queue.Queue()
            self._thread = threading.Thread(target=self._worker, daemon=True)
            self._thread.start()
        self._q.put((self._next, int(first), np.asarray(chain), np.asarray(log_prob), iteration, converged))
        self._next += 1

    def close(self):
        """Wait for the queued blocks to reach the disk."""
        if self._thread is not None:
            self._q.put(None)
            self._thread.join()
            self._thread = None
        if self._error is not None:
            err, self._error = self._error, None
            raise err
