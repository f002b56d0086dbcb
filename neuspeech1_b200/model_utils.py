"""Mirror of the reference's utils/model_utils.py stem factory (same name, arguments and error behaviour).

`projection_module('base', meg_ch=C, d_model=d)` returns the module tree the reference installs with
`encoder.set_input_embeddings` (utils/model_utils.py:9-17): Sequential(Conv1d(C,d,3,p1), GELU, Conv1d(d,d,3,s2,p1)) with the
metadata attribute `.stride = (2,)`.  Here the modules only carry the (trainable, fp32 master) parameters and their names
(`0.weight`, `2.weight`, ...); the arithmetic runs in the implicit-GEMM convolution kernels (ns_conv3_fwd/dgrad/wgrad)."""
import torch.nn as nn


def projection_module(config_name='', **kwargs):
    if config_name == 'base':
        d_model = kwargs['d_model']
        conv1 = nn.Sequential(
            nn.Conv1d(kwargs['meg_ch'], d_model, kernel_size=3, padding=1),
            nn.GELU(),
            nn.Conv1d(d_model, d_model, kernel_size=3, stride=2, padding=1),
        )
        conv1.stride = (2,)
    elif config_name == 'replace':
        # single strided conv (utils/model_utils.py:18-20): produces 750 positions for a 6000-sample input, which the
        # reference itself cannot run (utils/load_model.py:415 documents the shape failure); not on the hot path.
        raise NotImplementedError("projection_module('replace') is not supported by the B200 engine")
    else:
        raise NotImplementedError
    return conv1
