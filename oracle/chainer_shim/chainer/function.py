"""chainer.function.Function: forward(inputs) / backward(inputs, grad_outputs) on raw arrays."""
from .variable import Variable


class Function(object):
    def __call__(self, *inputs):
        inputs = [x if isinstance(x, Variable) else Variable(x) for x in inputs]
        outs = self.forward(tuple(x.data for x in inputs))
        if not isinstance(outs, tuple):
            outs = (outs,)
        self.inputs = inputs
        self.rank = max([x.rank for x in inputs] + [0])
        self.outputs = []
        for o in outs:
            v = Variable(o)
            v.creator = self
            v.rank = self.rank + 1
            self.outputs.append(v)
        return self.outputs[0] if len(self.outputs) == 1 else tuple(self.outputs)

    def check_type_forward(self, in_types):
        pass

    def forward(self, inputs):
        raise NotImplementedError

    def backward(self, inputs, grad_outputs):
        raise NotImplementedError
