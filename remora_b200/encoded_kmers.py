"""GPU replacement for the reference's Cython op ``remora.encoded_kmers.compute_encoded_kmer_batch``
(src/remora/encoded_kmers.pyx:13-45): same arguments, same float32 [n_chunks, 4*kmer_len, sig_len]
result (bit-exact), computed by the sm_100a kernel behind ``rb200_encode_dense``."""
import ctypes

import numpy as np
import torch

from . import RemoraError, _native


def compute_encoded_kmer_batch_torch(before_context_bases, after_context_bases, seqs, seq_mappings,
                                     seq_lens, sig_len=None, device=None):
    """Device tensors in, device tensor out.  ``sig_len`` defaults to the reference's rule
    ``seq_mappings[0, seq_lens[0]]`` (pyx:23), which costs one device->host read."""
    lib = _native.load_library()
    if device is None:
        device = seqs.device if isinstance(seqs, torch.Tensor) and seqs.is_cuda else \
            torch.device("cuda", torch.cuda.current_device())
    seqs = torch.as_tensor(seqs, dtype=torch.int8).to(device).contiguous()
    maps = torch.as_tensor(seq_mappings, dtype=torch.int16).to(device).contiguous()
    lens = torch.as_tensor(seq_lens, dtype=torch.int16).to(device).contiguous()
    n = lens.shape[0]
    kmer_len = before_context_bases + after_context_bases + 1
    if sig_len is None:
        if n == 0:
            raise RemoraError("cannot derive sig_len from an empty batch")
        sig_len = int(maps[0, int(lens[0])])
    out = torch.empty((n, 4 * kmer_len, sig_len), dtype=torch.float32, device=device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    with torch.cuda.device(device):
        rc = lib.rb200_encode_dense(ctypes.c_void_p(seqs.data_ptr()), seqs.shape[1],
                                    ctypes.c_void_p(maps.data_ptr()), maps.shape[1],
                                    ctypes.c_void_p(lens.data_ptr()), n, before_context_bases,
                                    after_context_bases, sig_len,
                                    ctypes.c_void_p(out.data_ptr()), stream)
    _native.check(rc, "rb200_encode_dense")
    return out


def compute_encoded_kmer_batch(before_context_bases, after_context_bases, seqs, seq_mappings,
                               seq_lens):
    """numpy in / numpy out, drop-in for the Cython function."""
    seqs = np.ascontiguousarray(seqs, dtype=np.int8)
    seq_mappings = np.ascontiguousarray(seq_mappings, dtype=np.int16)
    seq_lens = np.ascontiguousarray(seq_lens, dtype=np.int16)
    sig_len = int(seq_mappings[0, seq_lens[0]])
    out = compute_encoded_kmer_batch_torch(
        before_context_bases, after_context_bases, torch.from_numpy(seqs),
        torch.from_numpy(seq_mappings), torch.from_numpy(seq_lens), sig_len=sig_len,
        device=torch.device("cuda", torch.cuda.current_device()))
    return out.cpu().numpy()
