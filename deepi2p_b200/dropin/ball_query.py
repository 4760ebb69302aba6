"""Drop-in for the reference's torch extension `ball_query` (models/ball_query_ext/ball_query.cpp:45-48).

forward_cuda is a stub in the reference (prints "Not implemented yet." and returns nothing,
ball_query.cpp:23-31); here both names run the same kernel.
"""
from deepi2p_b200.point_ops import ball_query_forward as _fwd


def forward_cuda(node_to_point_dist, radius, K):
    return _fwd(node_to_point_dist, radius, K)


def forward_cuda_shared_mem(node_to_point_dist, radius, K):
    return _fwd(node_to_point_dist, radius, K)
