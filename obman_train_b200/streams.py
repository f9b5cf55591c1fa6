"""A second CUDA stream for work that is off the critical path of a training step.

The backward pass of the encoder / decoder is a serial chain of data-gradient kernels (dgrad of layer L feeds layer
L-1); the weight gradients, the BatchNorm-gradient reductions and the per-step weight folding hang off that chain
and nothing waits for them until the optimiser runs.  Issued on one stream they serialise with the chain (and every
small kernel adds its launch latency and tail to it); issued on this auxiliary stream they fill the SMs that the
chain's kernels leave idle (tiles of the last wave, prologues / epilogues, grids smaller than 148 CTAs).  The fork /
join pattern below is capturable, so the CUDA graph of the step keeps the same concurrency.

Memory rule (PyTorch caching allocator): a tensor that is produced on one stream and consumed on the other must stay
referenced until the two streams have been joined again - callers keep such tensors in a list until ``join()``.
``OBMAN_OVERLAP=0`` turns the auxiliary stream off (everything runs on the current stream).
"""
import contextlib
import os

import torch

_enabled = os.environ.get("OBMAN_OVERLAP", "1") != "0"
_streams = {}


def set_enabled(flag):
    """Returns the previous setting."""
    global _enabled
    prev, _enabled = _enabled, bool(flag)
    return prev


def enabled():
    return _enabled


WGRAD, CHAIN, BRANCH = 0, 1, 2   # auxiliary streams: weight-gradient work | second lane of the data-gradient chain
                                  # (phases of a strided dgrad, downsample branch) | AtlasNet branch next to MANO


def aux_stream(which=WGRAD):
    key = (torch.cuda.current_device(), which)
    if key not in _streams:
        _streams[key] = torch.cuda.Stream(device=key[0])
    return _streams[key]


def fork(which=WGRAD):
    """Order the auxiliary stream after everything issued so far on the current stream."""
    if _enabled:
        aux_stream(which).wait_stream(torch.cuda.current_stream())


def join(which=WGRAD):
    """Order the current stream after everything issued so far on the auxiliary stream."""
    if _enabled:
        torch.cuda.current_stream().wait_stream(aux_stream(which))


@contextlib.contextmanager
def on_aux(which=WGRAD):
    """Issue the enclosed launches on the auxiliary stream (no-op when overlap is disabled)."""
    if not _enabled:
        yield
        return
    with torch.cuda.stream(aux_stream(which)):
        yield
