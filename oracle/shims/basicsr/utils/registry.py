class Registry:
    """Minimal name -> object registry with the decorator form `@R.register()`."""

    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(fn_or_cls):
                self._obj_map[fn_or_cls.__name__] = fn_or_cls
                return fn_or_cls
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map

    def keys(self):
        return self._obj_map.keys()


ARCH_REGISTRY = Registry('arch')
DATASET_REGISTRY = Registry('dataset')
MODEL_REGISTRY = Registry('model')
LOSS_REGISTRY = Registry('loss')
METRIC_REGISTRY = Registry('metric')
