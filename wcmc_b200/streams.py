"""Two-stream execution of the diffuse / specular halves of the step.

KPCN's two branches and the two path-embedding networks are independent until the radiance
recombination (sbmc.KPCN.forward) / the loss (support/interfaces.py:206-251).  Every heavy kernel here is a
persistent grid of <= 148 CTAs whose last wave is partial, and the deep U-Net levels launch fewer CTAs than
there are SMs; with the halves on two streams the idle SMs of one kernel's tail run the other half's next
kernel.  CUDA graphs capture the fork / join as a DAG.

    with streams.fork("diffuse"):  ...work of one half...
    with streams.fork("specular"): ...work of the other half...
    streams.join()                 # the caller's stream waits for both

Memory safety: a side stream always starts by waiting for the caller's stream (so blocks of its pool that
were last read on the caller's stream are free to reuse), and join() makes the caller's stream wait for the
side streams before anything produced there is consumed.  WCMC_BRANCH_STREAMS=0 runs everything on the
caller's stream.
"""
import contextlib
import os
import threading

import torch

ENABLED = os.environ.get("WCMC_BRANCH_STREAMS", "1") != "0"
_side = {}      # (device index, name, thread) -> torch.cuda.Stream
_tls = threading.local()   # .open: side streams forked by THIS thread since its last join() (nn.DataParallel: one
                           # replica per thread, each with its own fork / join bookkeeping)


def _open():
    lst = getattr(_tls, "open", None)
    if lst is None:
        lst = _tls.open = []
    return lst


@contextlib.contextmanager
def fork(name):
    if not ENABLED or not torch.cuda.is_available():
        yield
        return
    dev = torch.cuda.current_device()
    cur = torch.cuda.current_stream(dev)
    key = (dev, name, threading.get_ident())
    st = _side.get(key)
    if st is None:
        st = _side[key] = torch.cuda.Stream(dev)
    if st == cur:           # nested use (KPCN inside a forked region): stay on the current stream
        yield
        return
    st.wait_stream(cur)
    _open().append((cur, st))
    with torch.cuda.stream(st):
        yield


def share():
    """How many such launches run side by side: 2 on a side stream of a fork (the other half's kernels are in flight on
    its sibling -- also during the backward pass, which autograd runs on the forward pass's streams), else 1.  The
    persistent convolution launches then plan for half of the SMs each (lib.conv2d, flags bits 24..27), so that the two
    halves overlap each other's prologue / tail / partial last wave instead of taking the whole machine in turns."""
    if not ENABLED or not _side or not torch.cuda.is_available():
        return 1
    cur = torch.cuda.current_stream()
    return 2 if any(cur == st for st in _side.values()) else 1


def join():
    lst = _open()
    while lst:
        cur, st = lst.pop()
        cur.wait_stream(st)
