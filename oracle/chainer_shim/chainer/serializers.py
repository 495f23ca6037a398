"""chainer.serializers: HDF5 is absent here; the reference's save/load are not exercised by the fixtures."""


def save_hdf5(filename, target):
    raise NotImplementedError("h5py is absent")


def load_hdf5(filename, target):
    raise NotImplementedError("h5py is absent")
