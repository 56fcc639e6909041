"""Dispatch to the tensor-core (tcgen05, NLC, bf16) pipeline when a forward call is eligible.
Returns None when it is not, and the caller continues on the generic NCL path."""


def try_wavenet_forward(model, signal):
    return None
