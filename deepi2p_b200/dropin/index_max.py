"""Drop-in for the reference's torch extension `index_max` (models/index_max_ext/index_max.cpp:154-159).

forward_cuda and forward_cuda_shared_mem run the sm_100a segmented-argmax kernel.  The reference's
CPU entry points (forward_cpu, forward_multi_thread_cpu) are outside the GPU hot path and are NOT
provided: this framework has no CPU compute path, and calling them raises.
"""
from deepi2p_b200.point_ops import index_max_forward as _fwd


def forward_cuda(data, index, K):
    return _fwd(data, index, K)


def forward_cuda_shared_mem(data, index, K):
    return _fwd(data, index, K)


def forward_cpu(data, index, K):
    raise NotImplementedError("deepi2p_b200 has no CPU path; use forward_cuda / forward_cuda_shared_mem")


def forward_multi_thread_cpu(data, index, K, thread_num):
    raise NotImplementedError("deepi2p_b200 has no CPU path; use forward_cuda / forward_cuda_shared_mem")
