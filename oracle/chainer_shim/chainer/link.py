"""chainer.Link / chainer.Chain: parameter registry in registration order."""
from .variable import Variable


class Link(object):
    def __init__(self):
        self._params = []

    def add_param(self, name, array):
        v = Variable(array, name=name)
        self._params.append(name)
        setattr(self, name, v)

    def params(self):
        for name in self._params:
            yield getattr(self, name)

    def namedparams(self):
        for name in self._params:
            yield "/" + name, getattr(self, name)

    def cleargrads(self):
        for p in self.params():
            p.cleargrad()

    def to_gpu(self):
        return self

    def to_cpu(self):
        return self


class Chain(Link):
    def __init__(self):
        Link.__init__(self)
        self._children = []

    def add_link(self, name, link):
        self._children.append(name)
        setattr(self, name, link)

    def params(self):
        for name in self._children:
            for p in getattr(self, name).params():
                yield p

    def namedparams(self):
        for name in self._children:
            for path, p in getattr(self, name).namedparams():
                yield "/" + name + path, p
