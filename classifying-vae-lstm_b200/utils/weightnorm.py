"""
Optimizer description objects: re-host of utils/weightnorm.py.  The arithmetic of
AdamWithWeightnorm.get_updates (utils/weightnorm.py:75-143) runs in the fused CUDA kernel
clv_adamwn_step; this class only carries the hyper-parameters the way the Keras optimizer object did.
SGDWithWeightnorm is never selected by any CLI and data_based_init is a no-op under Keras 2.0.0
(SURVEY quirk Q4): both are kept as explicit stubs.
"""


class AdamWithWeightnorm(object):
    name = "adam-wn"

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8, decay=0.0):
        if decay:
            raise NotImplementedError("decay is always 0.0 in the reference (utils/model_utils.py:54)")
        self.lr, self.beta_1, self.beta_2, self.epsilon, self.decay = lr, beta_1, beta_2, epsilon, decay


class Adam(AdamWithWeightnorm):
    name = "adam"


class SGDWithWeightnorm(object):
    def __init__(self, *a, **kw):
        raise NotImplementedError("SGDWithWeightnorm is never selected by the reference CLIs")


def data_based_init(model, input):
    """utils/weightnorm.py:182-210 only touches layers with attributes `W` and `b`; Keras-2.0.0
    layers expose kernel/bias, so in the reference this loop body never runs (quirk Q4).  No-op."""
    return None
