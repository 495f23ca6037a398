"""chainer.utils.type_check: the reference only calls expect() inside check_type_forward, which the shim never runs."""


def expect(*args):
    pass
