"""The update rules of the TF 1.13 optimizers the reference's train step uses
(models/sparse_graph_model.py:227-260: tf.train.GradientDescentOptimizer / RMSPropOptimizer / AdamOptimizer), as torch optimizers.

torch.optim.RMSprop / Adam are NOT the same rules: TF's RMSProp starts its mean-square slot at ONE and puts epsilon
inside the square root -- lr * g / sqrt(ms + eps) -- where torch starts at zero and uses sqrt(ms) + eps (early steps differ by up to
~7x); TF's Adam folds the bias corrections into the step size and uses the "epsilon hat" form.  A variable without a gradient
(an edge type absent from the batch) receives a ZERO gradient in TF, i.e. its slots still decay; `missing_grad_is_zero` mirrors that.
"""
import math

import torch


class TF1RMSProp(torch.optim.Optimizer):
    """tf.train.RMSPropOptimizer(learning_rate, decay, momentum, epsilon=1e-10), centered=False:
        ms  <- decay * ms + (1 - decay) * g^2            (ms initialised to ones)
        mom <- momentum * mom + lr * g / sqrt(ms + epsilon)
        var <- var - mom"""

    def __init__(self, params, lr, decay=0.9, momentum=0.0, epsilon=1e-10, missing_grad_is_zero=True):
        super().__init__(params, dict(lr=lr, decay=decay, momentum=momentum, epsilon=epsilon))
        self.missing_grad_is_zero = missing_grad_is_zero

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            lr, decay, momentum, eps = group["lr"], group["decay"], group["momentum"], group["epsilon"]
            for p in group["params"]:
                if p.grad is None and not self.missing_grad_is_zero:
                    continue
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                st = self.state[p]
                if not st:
                    st["ms"] = torch.ones_like(p)
                    st["mom"] = torch.zeros_like(p)
                st["ms"].mul_(decay).addcmul_(g, g, value=1.0 - decay)
                st["mom"].mul_(momentum).add_(g / torch.sqrt(st["ms"] + eps), alpha=lr)
                p.sub_(st["mom"])


class TF1Adam(torch.optim.Optimizer):
    """tf.train.AdamOptimizer(learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        lr_t <- lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m <- beta1 m + (1 - beta1) g;  v <- beta2 v + (1 - beta2) g^2
        var  <- var - lr_t * m / (sqrt(v) + epsilon)"""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, missing_grad_is_zero=True):
        super().__init__(params, dict(lr=lr, beta1=beta1, beta2=beta2, epsilon=epsilon))
        self.missing_grad_is_zero = missing_grad_is_zero
        self._t = 0

    @torch.no_grad()
    def step(self):
        self._t += 1
        for group in self.param_groups:
            b1, b2, eps = group["beta1"], group["beta2"], group["epsilon"]
            lr_t = group["lr"] * math.sqrt(1.0 - b2 ** self._t) / (1.0 - b1 ** self._t)
            for p in group["params"]:
                if p.grad is None and not self.missing_grad_is_zero:
                    continue
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                st = self.state[p]
                if not st:
                    st["m"] = torch.zeros_like(p)
                    st["v"] = torch.zeros_like(p)
                st["m"].mul_(b1).add_(g, alpha=1.0 - b1)
                st["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
                p.addcdiv_(st["m"], torch.sqrt(st["v"]) + eps, value=-lr_t)
