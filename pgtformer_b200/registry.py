"""ARCH_REGISTRY: BasicSR's when basicsr is installed, else a minimal stand-in with the same
`register()` / `get()` surface (the reference leaves PGTFormer unregistered,
`archs/pgtformer_arch.py:489`; this repo registers it)."""


class _Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._obj_map[o.__name__] = o
                return o
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError("No object named '%s' found in '%s' registry!" % (name, self._name))
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map

    def keys(self):
        return self._obj_map.keys()


try:                                                    # pragma: no cover - basicsr absent in this image
    from basicsr.utils.registry import ARCH_REGISTRY    # type: ignore
except Exception:
    ARCH_REGISTRY = _Registry('arch')
