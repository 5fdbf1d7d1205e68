"""Import shim so the untouched reference `gaussian_process.py` imports without autograd (only `train()` needs it;
hyper-parameter training is out of scope). Test infrastructure only."""


def value_and_grad(fn):
    def _unavailable(*a, **k):
        raise NotImplementedError("autograd is not installed; hyper-parameter training is out of scope")
    return _unavailable
