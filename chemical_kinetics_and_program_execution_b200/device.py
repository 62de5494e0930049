"""Device-resident use of the library: torch tensors for memory and streams, the C ABI for the
work.  Importing this module imports `markov_tapes` (runtime initialisation + known-answer test).
"""

import ctypes

import numpy
import torch

from . import _lib
from . import markov_tapes


def _current_stream_handle():
  """cudaStream_t of torch's current stream.  Torch reports the legacy default stream as 0,
  which the C ABI reads as "use the model's own stream", so it is passed as cudaStreamLegacy."""
  return torch.cuda.current_stream().cuda_stream or 1


class DeviceModel:
  """The structure of (tag, cl_k) on the current CUDA device.

  part=(rank, world) builds this rank's share only: the flux rules of the problem are dealt to
  `world` ranks (tapes_model_part), for any registered problem; the sum of the parts' dy/dt is the
  dy/dt of the whole problem (parallel.PeerExchangeRhs forms it over NVLink)."""

  def __init__(self, tag, cl_k, part=None):
    self.tag, self.cl_k, self.part = tag, cl_k, part
    if part is None:
      self.handle = markov_tapes.u_lib.tapes_model(tag.encode(), cl_k)
    else:
      self.handle = markov_tapes.u_lib.tapes_model_part(tag.encode(), cl_k, int(part[0]), int(part[1]))
    _lib.check(bool(self.handle), 'tapes_model')
    self.info = _lib.model_info(self.handle)
    self.timing = _lib.model_timing(self.handle)
    self.n_states = self.info['n_states']

  def set_option(self, key, value):
    rc = markov_tapes.u_lib.tapes_model_set(self.handle, key.encode(), int(value))
    _lib.check(rc == 0, 'tapes_model_set')
    self.info = _lib.model_info(self.handle)

  def rhs(self, p, out=None):
    """dy/dt of a float64 CUDA tensor, asynchronously on torch's current stream."""
    assert p.is_cuda and p.dtype == torch.float64 and p.is_contiguous() and p.numel() == self.n_states
    if out is None:
      out = torch.empty_like(p)
    stream = _current_stream_handle()
    rc = markov_tapes.u_lib.tapes_rhs_device(self.handle, p.data_ptr(), out.data_ptr(), stream)
    _lib.check(rc == 0, 'tapes_rhs_device')
    return out

  def weights(self, p):
    """Everything of a right-hand side that depends on p (asynchronous, current stream)."""
    rc = markov_tapes.u_lib.tapes_weights_device(self.handle, p.data_ptr(), _current_stream_handle())
    _lib.check(rc == 0, 'tapes_weights_device')

  def flux_rows(self, out, row_lo, row_hi):
    """dy/dt of the states row_lo <= i < row_hi from the weights of the last `weights` call."""
    rc = markov_tapes.u_lib.tapes_flux_rows_device(self.handle, out.data_ptr(), int(row_lo), int(row_hi),
                                                   _current_stream_handle())
    _lib.check(rc == 0, 'tapes_flux_rows_device')

  def rhs_profile(self, p, out):
    """One right-hand side with CUDA events between its phases; returns ms per phase
    (marginals + leaf-world probabilities, forest levels, S*w) measured on the launch stream."""
    stream = _current_stream_handle()
    ms = numpy.zeros(3, dtype=numpy.float64)
    rc = markov_tapes.u_lib.tapes_rhs_profile(self.handle, p.data_ptr(), out.data_ptr(), stream,
                                              ms.ctypes.data, 3)
    _lib.check(rc == 0, 'tapes_rhs_profile')
    return ms

  def observe(self, y, seqs, eps=None):
    """Sequence probabilities of a CUDA table evaluated on the device: `seq_prob` of
    framework/markov_tapes.py:190-233 for sequences of any length."""
    assert y.is_cuda and y.dtype == torch.float64 and y.is_contiguous() and y.numel() == self.n_states
    torch.cuda.current_stream().synchronize()  # the sums run on the model's stream
    out = numpy.zeros(len(seqs), dtype=numpy.float64)
    if len(seqs):
      seq_ptr, symbols = _lib.pack_sequences(seqs)
      rc = markov_tapes.u_lib.tapes_observe_sequences(self.handle, y.data_ptr(), len(seqs), seq_ptr.ctypes.data,
                                                      symbols.ctypes.data, 1e-100 if eps is None else float(eps),
                                                      out.ctypes.data)
      _lib.check(rc == 0, 'tapes_observe_sequences')
    return out

  def entropy(self, y):
    """`markov_entropy` (framework/markov_tapes.py:178-187) of a CUDA table, evaluated on the device."""
    assert y.is_cuda and y.dtype == torch.float64 and y.is_contiguous() and y.numel() == self.n_states
    torch.cuda.current_stream().synchronize()
    out = numpy.zeros(1, dtype=numpy.float64)
    _lib.check(markov_tapes.u_lib.tapes_markov_entropy(self.handle, y.data_ptr(), out.ctypes.data) == 0,
               'tapes_markov_entropy')
    return float(out[0])

  def csr(self):
    """(row_ptr[int64], entries[uint32]) copied to host."""
    row_ptr = numpy.zeros(self.n_states + 1, dtype=numpy.int64)
    entries = numpy.zeros(max(self.info['nnz'], 1), dtype=numpy.uint32)
    rc = markov_tapes.u_lib.tapes_export_csr(self.handle, row_ptr.ctypes.data, entries.ctypes.data)
    _lib.check(rc == 0, 'tapes_export_csr')
    return row_ptr, entries[:self.info['nnz']]

  def node_weights(self):
    w = numpy.zeros(max(self.info['n_nodes'], 1), dtype=numpy.float64)
    rc = markov_tapes.u_lib.tapes_export_node_weights(self.handle, w.ctypes.data)
    _lib.check(rc == 0, 'tapes_export_node_weights')
    return w[:self.info['n_nodes']]

  def terms(self):
    """Flux terms of the most recent right-hand side as (src, dst, w) arrays in node order."""
    row_ptr, entries = self.csr()
    w = self.node_weights()
    rows = numpy.repeat(numpy.arange(self.n_states, dtype=numpy.int64), numpy.diff(row_ptr))
    node = (entries & 0x7fffffff).astype(numpy.int64)
    outflow = (entries >> 31).astype(bool)
    src = numpy.full(len(w), -1, dtype=numpy.int64)
    dst = numpy.full(len(w), -1, dtype=numpy.int64)
    src[node[outflow]] = rows[outflow]
    dst[node[~outflow]] = rows[~outflow]
    has = src >= 0
    assert (has == (dst >= 0)).all()
    return src[has], dst[has], w[has]
