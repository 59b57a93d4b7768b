"""QNetwork: the DDQN agent's Q-network (util/model/QNets.py:7-27,55-60, EvalModel / TargetModel) evaluated for a whole batch of
environments by one hand-written tcgen05 kernel (csrc/abx_qnet.cu) through the C ABI (abx_qnet_* in include/abides_b200.h).

    net = QNetwork()                                  # 2 -> 32 -> 64 -> 128 -> 128 -> 64 -> 32 -> 24, Glorot-uniform weights like Keras
    q, actions = net.forward(obs, x_offset=6)         # obs: CUDA fp64 [n, 8] from DDQNExecutionEnv.step; reads columns 6, 7

`actions` follow the agent's rule (ddqlearning_execution_agent.py:339-365): np.argmax of the Q row, or with `greedy_prob` < 1 the
epsilon branch (greedy with that probability, else uniform over the actions; Philox instead of np.random).  There is no CPU
fallback; `torch_reference` is the plain PyTorch fp32 network the tests compare against.
"""
import ctypes as C

import numpy as np

from . import _lib

DEFAULT_DIMS = (2, 32, 64, 128, 128, 64, 32, 24)      # the reference's state is 2-dimensional: discretize() zips 6 features with a 2-entry grid
NNMODEL_2_DIMS = (2, 32, 64, 128, 256, 128, 64, 32, 24)  # util/model/QNets.py:30-52 NNModel_2: too large for shared memory, evaluated by the streamed kernel (layer widths up to 256)


def param_count(dims):
    return sum(dims[l] * dims[l + 1] + dims[l + 1] for l in range(len(dims) - 1))


def init_params(dims=DEFAULT_DIMS, seed=0):
    """Keras Dense defaults: kernel glorot_uniform, bias zeros.  Flat fp32 vector: per layer W[out][in] (row major) then b[out]."""
    rs = np.random.RandomState(seed)
    parts = []
    for l in range(len(dims) - 1):
        lim = np.sqrt(6.0 / (dims[l] + dims[l + 1]))
        parts.append(rs.uniform(-lim, lim, size=(dims[l + 1], dims[l])).astype(np.float32).ravel())
        parts.append(np.zeros(dims[l + 1], dtype=np.float32))
    return np.concatenate(parts)


def unpack_params(flat, dims=DEFAULT_DIMS):
    out, p = [], 0
    for l in range(len(dims) - 1):
        n = dims[l] * dims[l + 1]
        w = np.asarray(flat[p:p + n], dtype=np.float32).reshape(dims[l + 1], dims[l])
        p += n
        b = np.asarray(flat[p:p + dims[l + 1]], dtype=np.float32)
        p += dims[l + 1]
        out.append((w, b))
    return out


def torch_reference(flat, x, dims=DEFAULT_DIMS):
    """Plain PyTorch fp32 forward of the same network (x: [n, dims[0]] tensor)."""
    import torch
    h = x.to(torch.float32)
    layers = unpack_params(np.asarray(flat), dims)
    for i, (w, b) in enumerate(layers):
        h = h @ torch.from_numpy(w).to(h.device).T + torch.from_numpy(b).to(h.device)
        if i + 1 < len(layers):
            h = torch.relu(h)
    return h


class QNetwork:
    def __init__(self, dims=DEFAULT_DIMS, params=None, device=0, seed=0, lib_path=None):
        self._L = _lib.load(lib_path)
        self.dims = tuple(int(d) for d in dims)
        self.n_layers = len(self.dims) - 1
        self.device = int(device)
        self.params = np.ascontiguousarray(init_params(self.dims, seed) if params is None else params, dtype=np.float32)
        if self.params.size != param_count(self.dims):
            raise ValueError("expected %d parameters, got %d" % (param_count(self.dims), self.params.size))
        d = (C.c_int32 * len(self.dims))(*self.dims)
        self._h = C.c_void_p()
        st = self._L.abx_qnet_create(d, self.n_layers, self.params.ctypes.data_as(C.POINTER(C.c_float)), self.device, C.byref(self._h))
        if st != 0:
            raise _lib.AbxError("abx_qnet_create failed: %s (%s)" % (self._L.abx_strerror(st).decode(), self._L.abx_qnet_last_error().decode()))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.abx_qnet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, flat, stream=None):
        """Replace the weights (the learner's update, or the target-network sync of train_neural_nets :486-490)."""
        self.params = np.ascontiguousarray(flat, dtype=np.float32)
        st = self._L.abx_qnet_set_params(self._h, self.params.ctypes.data_as(C.POINTER(C.c_float)), stream)
        if st != 0:
            raise _lib.AbxError("abx_qnet_set_params failed: %s" % self._L.abx_qnet_last_error().decode())

    def set_params_device(self, flat_dev, stream=None):
        """Same from a CUDA fp32 tensor, without leaving the device (asynchronous on the current stream); `self.params` (the host copy) is
        NOT refreshed -- call `sync_host_params` when the host copy is needed."""
        import torch
        if not (isinstance(flat_dev, torch.Tensor) and flat_dev.is_cuda and flat_dev.dtype == torch.float32 and flat_dev.is_contiguous()
                and flat_dev.numel() == param_count(self.dims)):
            raise ValueError("flat_dev must be a contiguous CUDA fp32 tensor with %d elements" % param_count(self.dims))
        sp = C.c_void_p(torch.cuda.current_stream(flat_dev.device).cuda_stream) if stream is None else stream
        st = self._L.abx_qnet_set_params_device(self._h, C.c_void_p(flat_dev.data_ptr()), sp)
        if st != 0:
            raise _lib.AbxError("abx_qnet_set_params_device failed: %s" % self._L.abx_qnet_last_error().decode())
        self._dev_params = flat_dev

    def sync_host_params(self):
        if getattr(self, "_dev_params", None) is not None:
            self.params = self._dev_params.detach().cpu().numpy().astype(np.float32)
        return self.params

    def forward(self, x, x_offset=0, want_q=True, want_actions=True, greedy_prob=1.0, seed=0, counter=0, out=None, stream=None):
        """x: CUDA fp64 tensor [n, stride]; the state is x[:, x_offset : x_offset + dims[0]].  Returns (q fp32 [n, n_out] | None,
        actions int32 [n] | None) as CUDA tensors."""
        import torch
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.is_contiguous()):
            raise ValueError("x must be a contiguous CUDA fp64 tensor [n, stride]")
        n, stride = x.shape
        if out is not None:
            q, act = out
        else:
            q = torch.empty(n, self.dims[-1], dtype=torch.float32, device=x.device) if want_q else None
            act = torch.empty(n, dtype=torch.int32, device=x.device) if want_actions else None
        sp = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream) if stream is None else stream
        st = self._L.abx_qnet_forward(self._h, C.c_void_p(x.data_ptr()), int(stride), int(x_offset), int(n),
                                      C.c_void_p(q.data_ptr()) if q is not None else None, C.c_void_p(act.data_ptr()) if act is not None else None,
                                      C.c_double(greedy_prob), C.c_uint64(seed), C.c_uint64(counter), sp)
        if st != 0:
            raise _lib.AbxError("abx_qnet_forward failed: %s (%s)" % (self._L.abx_strerror(st).decode(), self._L.abx_qnet_last_error().decode()))
        return q, act

    @property
    def launch_count(self):
        return int(self._L.abx_qnet_launch_count(self._h))
