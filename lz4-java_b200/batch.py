"""Batch entry points: one launch over n independent blocks (SURVEY.md §7 hard part 1, §8e).

Host variants take numpy arrays (pinned or not) and run the library's H2D / kernel / D2H pipeline;
device variants take torch CUDA tensors that are already resident in HBM and launch on torch's
current stream.  torch is plumbing here (device memory + streams); it is imported lazily so the
host API works without it.
"""
from __future__ import annotations

import numpy as np

from . import _native as N


def _np(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def _p(a) -> int:
    return a.ctypes.data


def uniform_layout(n: int, block: int, stride: int | None = None):
    """offsets/lengths for n equal blocks laid out at `stride` (default: back to back)."""
    stride = block if stride is None else stride
    off = np.arange(n, dtype=np.uint64) * np.uint64(stride)
    return off, np.full(n, block, dtype=np.int32)


# ------------------------------------------------------------------ host buffers
def compress_fast_batch_host(src, src_off, src_len, dst, dst_off, dst_cap, max_src_len=0) -> np.ndarray:
    src_off, src_len, dst_off, dst_cap = _np(src_off, np.uint64), _np(src_len, np.int32), _np(dst_off, np.uint64), _np(dst_cap, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    N.check(N.lib().b200lz4_compress_fast_batch_host(_p(src), _p(src_off), _p(src_len), _p(dst), _p(dst_off), _p(dst_cap),
                                                     _p(res), len(src_off), max_src_len))
    return res


def compress_hc_batch_host(src, src_off, src_len, dst, dst_off, dst_cap, level=9) -> np.ndarray:
    src_off, src_len, dst_off, dst_cap = _np(src_off, np.uint64), _np(src_len, np.int32), _np(dst_off, np.uint64), _np(dst_cap, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    N.check(N.lib().b200lz4_compress_hc_batch_host(_p(src), _p(src_off), _p(src_len), _p(dst), _p(dst_off), _p(dst_cap),
                                                   _p(res), len(src_off), level))
    return res


def compress_fast_compact_host(src, src_off, src_len, dst, max_src_len=0):
    """-> (out_off[u64], out_len[i32], total)"""
    src_off, src_len = _np(src_off, np.uint64), _np(src_len, np.int32)
    n = len(src_off)
    out_off = np.zeros(n, dtype=np.uint64)
    res = np.zeros(n, dtype=np.int32)
    total = np.zeros(1, dtype=np.uint64)
    N.check(N.lib().b200lz4_compress_fast_compact_host(_p(src), _p(src_off), _p(src_len), _p(dst), dst.nbytes, _p(out_off),
                                                       _p(res), n, max_src_len, _p(total)))
    return out_off, res, int(total[0])


def decompress_safe_batch_host(src, src_off, src_len, dst, dst_off, dst_cap) -> np.ndarray:
    src_off, src_len, dst_off, dst_cap = _np(src_off, np.uint64), _np(src_len, np.int32), _np(dst_off, np.uint64), _np(dst_cap, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    N.check(N.lib().b200lz4_decompress_safe_batch_host(_p(src), _p(src_off), _p(src_len), _p(dst), _p(dst_off), _p(dst_cap),
                                                       _p(res), len(src_off)))
    return res


def decompress_fast_batch_host(src, src_off, src_avail, dst, dst_off, dst_len) -> np.ndarray:
    src_off, src_avail, dst_off, dst_len = _np(src_off, np.uint64), _np(src_avail, np.int32), _np(dst_off, np.uint64), _np(dst_len, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    N.check(N.lib().b200lz4_decompress_fast_batch_host(_p(src), _p(src_off), _p(src_avail), _p(dst), _p(dst_off), _p(dst_len),
                                                       _p(res), len(src_off)))
    return res


def xxh32_batch_host(buf, off, length, seed=0) -> np.ndarray:
    off, length = _np(off, np.uint64), _np(length, np.int32)
    out = np.zeros(len(off), dtype=np.uint32)
    N.check(N.lib().b200xxh32_batch_host(_p(buf), _p(off), _p(length), seed & 0xFFFFFFFF, _p(out), len(off)))
    return out


def xxh64_batch_host(buf, off, length, seed=0) -> np.ndarray:
    off, length = _np(off, np.uint64), _np(length, np.int32)
    out = np.zeros(len(off), dtype=np.uint64)
    N.check(N.lib().b200xxh64_batch_host(_p(buf), _p(off), _p(length), seed & 0xFFFFFFFFFFFFFFFF, _p(out), len(off)))
    return out


# ------------------------------------------------------------------ host buffers, one process driving several GPUs
def _devs(devices):
    """devices: an int (0..k-1) or an explicit list of device indices -> (ctypes int array or None, count)"""
    import ctypes
    if isinstance(devices, int):
        return None, devices
    arr = (ctypes.c_int * len(devices))(*devices)
    return arr, len(devices)


def compress_fast_batch_host_multi(src, src_off, src_len, dst, dst_off, dst_cap, devices, max_src_len=0) -> np.ndarray:
    """range-shards the batch over `devices` inside this process (b200lz4_compress_fast_batch_host_multi)"""
    src_off, src_len, dst_off, dst_cap = _np(src_off, np.uint64), _np(src_len, np.int32), _np(dst_off, np.uint64), _np(dst_cap, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    arr, k = _devs(devices)
    N.check(N.lib().b200lz4_compress_fast_batch_host_multi(_p(src), _p(src_off), _p(src_len), _p(dst), _p(dst_off), _p(dst_cap),
                                                           _p(res), len(src_off), max_src_len, arr, k))
    return res


def compress_fast_compact_host_multi(src, src_off, src_len, dst, devices, max_src_len=0):
    """-> (out_off[u64] absolute in dst, out_len[i32], shard_base[u64], shard_total[u64]): packed within each GPU's shard"""
    src_off, src_len = _np(src_off, np.uint64), _np(src_len, np.int32)
    n = len(src_off)
    out_off = np.zeros(n, dtype=np.uint64)
    res = np.zeros(n, dtype=np.int32)
    arr, k = _devs(devices)
    sbase, stotal = np.zeros(max(k, 1), dtype=np.uint64), np.zeros(max(k, 1), dtype=np.uint64)
    N.check(N.lib().b200lz4_compress_fast_compact_host_multi(_p(src), _p(src_off), _p(src_len), _p(dst), dst.nbytes, _p(out_off),
                                                             _p(res), n, max_src_len, arr, k, _p(sbase), _p(stotal)))
    return out_off, res, sbase, stotal


def compress_hc_batch_host_multi(src, src_off, src_len, dst, dst_off, dst_cap, devices, level=9) -> np.ndarray:
    src_off, src_len, dst_off, dst_cap = _np(src_off, np.uint64), _np(src_len, np.int32), _np(dst_off, np.uint64), _np(dst_cap, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    arr, k = _devs(devices)
    N.check(N.lib().b200lz4_compress_hc_batch_host_multi(_p(src), _p(src_off), _p(src_len), _p(dst), _p(dst_off), _p(dst_cap),
                                                         _p(res), len(src_off), level, arr, k))
    return res


def decompress_safe_batch_host_multi(src, src_off, src_len, dst, dst_off, dst_cap, devices) -> np.ndarray:
    src_off, src_len, dst_off, dst_cap = _np(src_off, np.uint64), _np(src_len, np.int32), _np(dst_off, np.uint64), _np(dst_cap, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    arr, k = _devs(devices)
    N.check(N.lib().b200lz4_decompress_safe_batch_host_multi(_p(src), _p(src_off), _p(src_len), _p(dst), _p(dst_off), _p(dst_cap),
                                                             _p(res), len(src_off), arr, k))
    return res


def decompress_fast_batch_host_multi(src, src_off, src_avail, dst, dst_off, dst_len, devices) -> np.ndarray:
    src_off, src_avail, dst_off, dst_len = _np(src_off, np.uint64), _np(src_avail, np.int32), _np(dst_off, np.uint64), _np(dst_len, np.int32)
    res = np.zeros(len(src_off), dtype=np.int32)
    arr, k = _devs(devices)
    N.check(N.lib().b200lz4_decompress_fast_batch_host_multi(_p(src), _p(src_off), _p(src_avail), _p(dst), _p(dst_off), _p(dst_len),
                                                             _p(res), len(src_off), arr, k))
    return res


def xxh32_batch_host_multi(buf, off, length, devices, seed=0) -> np.ndarray:
    off, length = _np(off, np.uint64), _np(length, np.int32)
    out = np.zeros(len(off), dtype=np.uint32)
    arr, k = _devs(devices)
    N.check(N.lib().b200xxh32_batch_host_multi(_p(buf), _p(off), _p(length), seed & 0xFFFFFFFF, _p(out), len(off), arr, k))
    return out


def xxh64_batch_host_multi(buf, off, length, devices, seed=0) -> np.ndarray:
    off, length = _np(off, np.uint64), _np(length, np.int32)
    out = np.zeros(len(off), dtype=np.uint64)
    arr, k = _devs(devices)
    N.check(N.lib().b200xxh64_batch_host_multi(_p(buf), _p(off), _p(length), seed & 0xFFFFFFFFFFFFFFFF, _p(out), len(off), arr, k))
    return out


# ------------------------------------------------------------------ device-resident (torch tensors)
def _stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def _dev_batch(fn, src, src_off, src_len, dst, dst_off, dst_cap, result, *extra):
    n = src_off.numel()
    N.check(fn(src.data_ptr(), src_off.data_ptr(), src_len.data_ptr(), dst.data_ptr(), dst_off.data_ptr(),
               dst_cap.data_ptr(), result.data_ptr(), n, *extra, _stream_ptr()))
    return result


def compress_fast_batch_dev(src, src_off, src_len, dst, dst_off, dst_cap, result, max_src_len=0):
    return _dev_batch(N.lib().b200lz4_compress_fast_batch_dev, src, src_off, src_len, dst, dst_off, dst_cap, result, max_src_len)


def compress_hc_batch_dev(src, src_off, src_len, dst, dst_off, dst_cap, result, level=9):
    return _dev_batch(N.lib().b200lz4_compress_hc_batch_dev, src, src_off, src_len, dst, dst_off, dst_cap, result, level)


def decompress_safe_batch_dev(src, src_off, src_len, dst, dst_off, dst_cap, result):
    return _dev_batch(N.lib().b200lz4_decompress_safe_batch_dev, src, src_off, src_len, dst, dst_off, dst_cap, result)


def decompress_fast_batch_dev(src, src_off, src_avail, dst, dst_off, dst_len, result):
    return _dev_batch(N.lib().b200lz4_decompress_fast_batch_dev, src, src_off, src_avail, dst, dst_off, dst_len, result)


def xxh32_batch_dev(buf, off, length, out, seed=0):
    N.check(N.lib().b200xxh32_batch_dev(buf.data_ptr(), off.data_ptr(), length.data_ptr(), seed & 0xFFFFFFFF,
                                        out.data_ptr(), off.numel(), _stream_ptr()))
    return out


def xxh64_batch_dev(buf, off, length, out, seed=0):
    N.check(N.lib().b200xxh64_batch_dev(buf.data_ptr(), off.data_ptr(), length.data_ptr(), seed & 0xFFFFFFFFFFFFFFFF,
                                        out.data_ptr(), off.numel(), _stream_ptr()))
    return out


def compact_dev(slots, slot_off, lens, out, out_off, total):
    """device-resident packing (b200lz4_compact_dev): out_off <- exclusive prefix sums of lens, total[0] <- their sum,
    block i's bytes move from slots + slot_off[i] to out + out_off[i].  Torch CUDA tensors, torch's current stream."""
    N.check(N.lib().b200lz4_compact_dev(slots.data_ptr(), slot_off.data_ptr(), lens.data_ptr(), out.data_ptr(), out_off.data_ptr(),
                                        total.data_ptr(), slot_off.numel(), _stream_ptr()))
    return out_off, total


def stitch_shards_dev(shards, totals, dst):
    """one process, several GPUs: packed shard g (a uint8 CUDA tensor on its own GPU, totals[g] bytes of it) lands in `dst`
    (a uint8 CUDA tensor on any GPU) at the sum of the totals before it -- peer copies, b200lz4_stitch_shards_dev.
    -> positions of the shards in dst (uint64)"""
    import ctypes
    import torch
    k = len(shards)
    for t in shards:
        torch.cuda.synchronize(t.device)                          # the shards must be complete (see the header)
    ptrs = (ctypes.c_void_p * k)(*[int(t.data_ptr()) for t in shards])
    devs = (ctypes.c_int * k)(*[int(t.device.index or 0) for t in shards])
    tot = np.asarray([int(x) for x in totals], dtype=np.uint64)
    pos = np.zeros(k, dtype=np.uint64)
    N.check(N.lib().b200lz4_stitch_shards_dev(ptrs, devs, _p(tot), k, dst.data_ptr(), int(dst.device.index or 0), dst.numel(), _p(pos)))
    return pos

