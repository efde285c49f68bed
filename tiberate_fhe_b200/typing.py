"""Layout-compatible data structures for the hot path (mirror of tiberate/typing.py:24-49 FLAGS,
:52-301 DataStruct, :417-667 ciphertext classes, :675-713 key classes).

Only the layout contract is reproduced: `data` of a Ciphertext is [c0, c1], each a
per-device list of int64 [limbs, N] tensors (a leading batch dimension is additionally accepted by
the B200 engine); a KeySwitchKey's data is a list over global digit-group id of PublicKey objects
whose data is [b_list, a_list] with [P, N] tensors.  Operator sugar, pickling and device moves of
the reference classes are host conveniences outside the hot path.
"""

from __future__ import annotations

from collections import defaultdict
from enum import Flag, auto


class FLAGS(Flag):
    NTT_STATE = auto()
    MONTGOMERY_STATE = auto()
    INCLUDE_SPECIAL = auto()
    NEED_RESCALE = auto()
    NEED_RELINERIZE = auto()


def _none():
    return None


class DataStruct:
    def __init__(self, data=None, *, flags=None, level: int, **kwargs):
        self.data = data
        self._flags = FLAGS(0)
        flags = flags or []
        if isinstance(flags, list):
            for f in flags:
                self._flags |= f
        elif isinstance(flags, FLAGS):
            self._flags = flags
        self.level = level
        self.misc = defaultdict(_none)
        if "misc" in kwargs:
            self.misc.update(kwargs.pop("misc") or {})
        self.misc.update(kwargs)

    def has_flag(self, flag: FLAGS) -> bool:
        return bool(self._flags & flag)

    def set_flag(self, flag: FLAGS):
        self._flags |= flag

    def rm_flag(self, flag: FLAGS):
        self._flags &= ~flag

    def clone(self, clone_data: bool = True):
        def cp(x):
            if isinstance(x, list):
                return [cp(y) for y in x]
            return x.clone() if hasattr(x, "clone") else x

        new = self.__class__(cp(self.data) if clone_data else [], flags=self._flags, level=self.level,
                             misc=dict(self.misc))
        return new

    # ---- on-disk format --------------------------------------------------------------
    # The reference pickles the Python object (tiberate/typing.py:283-290), i.e. loading a file executes
    # whatever the pickle says.  Here a file is a torch.save of plain containers -- class name, flags,
    # level, metadata and the nested tensor lists -- readable with `weights_only=True`.
    def _to_plain(self):
        def pack(x):
            if isinstance(x, DataStruct):
                return {"__tb200__": x._to_plain()}
            if isinstance(x, (list, tuple)):
                return [pack(y) for y in x]
            return x

        misc = {k: v for k, v in self.misc.items() if isinstance(v, (int, float, str, bool, type(None)))}
        return {"format": "tb200-datastruct-1", "cls": self.__class__.__name__, "flags": int(self._flags.value),
                "level": int(self.level), "misc": misc, "data": pack(self.data)}

    @classmethod
    def _from_plain(cls, d):
        def unpack(x):
            if isinstance(x, dict) and "__tb200__" in x:
                return DataStruct._from_plain(x["__tb200__"])
            if isinstance(x, list):
                return [unpack(y) for y in x]
            return x

        if d.get("format") != "tb200-datastruct-1":
            raise ValueError("not a tb200 data-structure file")
        klass = _CLASSES.get(d["cls"])
        if klass is None:
            raise ValueError(f"unknown data-structure class {d['cls']!r}")
        obj = klass.__new__(klass)
        DataStruct.__init__(obj, unpack(d["data"]), flags=FLAGS(d["flags"]), level=d["level"], misc=d["misc"])
        return obj

    def save(self, path: str):
        import torch

        torch.save(self._to_plain(), path)

    @classmethod
    def load(cls, path: str, map_location=None):
        import torch

        obj = DataStruct._from_plain(torch.load(path, map_location=map_location, weights_only=True))
        if cls is not DataStruct and not isinstance(obj, cls):
            raise TypeError(f"{path} holds a {type(obj).__name__}, not a {cls.__name__}")
        return obj

    @classmethod
    def wrap(cls, another: "DataStruct", **kwargs):
        return cls(another.data, flags=another._flags, level=another.level, misc=dict(another.misc), **kwargs)


class Ciphertext(DataStruct):
    pass


class CiphertextTriplet(DataStruct):
    pass


class Plaintext(DataStruct):
    """Holds the per-level cache of the NTT+Montgomery encoding used by pc_mult
    (tiberate/typing.py:318-409).  Encoding itself (float FFT) is outside the hot path: build one
    with `Plaintext.from_ntt(level, tensor)`."""

    def __init__(self, m=None, *, flags=None, level: int = 0, scale=None, **kwargs):
        """m: the 1-D message (tensor / array / list / scalar) as in the reference's Plaintext(m), or None
        when only a cache is supplied (from_ntt)."""
        super().__init__(None, flags=flags, level=level, **kwargs)
        if m is not None:
            import numpy as np
            import torch

            if isinstance(m, np.ndarray):
                m = torch.from_numpy(m)
            elif isinstance(m, (int, float)):
                m = torch.tensor([m])
            elif not isinstance(m, torch.Tensor):
                m = torch.tensor(m)
            if m.dim() != 1:
                raise RuntimeError(f"Plaintext source data must be 1D tensor, got {m.dim()}D tensor.")
        self.src = m
        self.scale = scale
        self.cache = defaultdict(dict)

    @classmethod
    def from_ntt(cls, level: int, pt_ntt):
        pt = cls(None, level=level)
        pt.cache[level]["pc_mult"] = pt_ntt if isinstance(pt_ntt, list) else [pt_ntt]
        return pt


class SecretKey(DataStruct):
    pass


class EvaluationKey(SecretKey):
    pass


class PublicKey(DataStruct):
    pass


class KeySwitchKey(DataStruct):
    pass


class RotationKey(KeySwitchKey):
    @property
    def delta(self):
        return self.misc.get("delta")

    @delta.setter
    def delta(self, value):
        self.misc["delta"] = value


class ConjugationKey(DataStruct):
    pass


_CLASSES = {c.__name__: c for c in (Ciphertext, CiphertextTriplet, SecretKey, EvaluationKey, PublicKey, KeySwitchKey,
                                    RotationKey, ConjugationKey)}


# ---- reading files written by the reference ------------------------------------------
# The reference saves a DataStruct with pickle.dump(self) (tiberate/typing.py:283-290).  Loading such a
# file with pickle.load would execute whatever the file says; this loader accepts exactly what a genuine
# file contains and nothing else: the tiberate.typing classes (mapped onto the classes of this module),
# defaultdict / OrderedDict, torch's tensor rebuild helpers, numpy array reconstruction, and the storage
# loader -- re-implemented on top of torch.load(weights_only=True).  Any other global raises UnpicklingError.
_REF_TYPING = {"Ciphertext": Ciphertext, "CiphertextTriplet": CiphertextTriplet, "Plaintext": Plaintext,
               "SecretKey": SecretKey, "EvaluationKey": EvaluationKey, "PublicKey": PublicKey,
               "KeySwitchKey": KeySwitchKey, "RotationKey": RotationKey, "GaloisKey": KeySwitchKey,
               "ConjugationKey": ConjugationKey, "DataStruct": DataStruct, "FLAGS": FLAGS, "_default_none": _none}


class _RefCachedDict(dict):
    """Stands in for vdtoys.cache.CachedDict (a dict with a generator function): only the stored items survive."""

    def __setstate__(self, state):
        if isinstance(state, dict):
            inner = state.get("_cache")
            if isinstance(inner, dict):
                self.update(inner)


def load_reference_pickle(path: str, map_location=None):
    """Ciphertext / key / plaintext file written by tiberate's DataStruct.save -> the matching class of this
    module (same `data` nesting, flags, level, misc), without executing code from the file."""
    import io
    import pickle

    import torch

    def safe_storage_from_bytes(b):
        return torch.load(io.BytesIO(b), weights_only=True, map_location=map_location)

    allowed = {
        ("collections", "defaultdict"): defaultdict,
        ("collections", "OrderedDict"): __import__("collections").OrderedDict,
        ("torch._utils", "_rebuild_tensor_v2"): torch._utils._rebuild_tensor_v2,
        ("torch._utils", "_rebuild_parameter"): torch._utils._rebuild_parameter,
        ("torch.storage", "_load_from_bytes"): safe_storage_from_bytes,
        ("torch", "device"): torch.device,
        ("torch", "Size"): torch.Size,
        ("vdtoys.cache", "CachedDict"): _RefCachedDict,
    }
    for name in ("int64", "int32", "float64", "float32", "complex128", "complex64", "bool", "uint8"):
        allowed[("torch", name)] = getattr(torch, name)

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module == "tiberate.typing" and name in _REF_TYPING:
                return _REF_TYPING[name]
            if (module, name) in allowed:
                return allowed[(module, name)]
            if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name in ("_reconstruct", "scalar"):
                import numpy.core.multiarray as m

                return getattr(m, name)
            if module == "numpy" and name in ("ndarray", "dtype"):
                import numpy

                return getattr(numpy, name)
            raise pickle.UnpicklingError(f"{path}: global {module}.{name} is not part of a tiberate data file")

    with open(path, "rb") as f:
        obj = Unpickler(f).load()
    if not isinstance(obj, DataStruct):
        raise pickle.UnpicklingError(f"{path} does not hold a tiberate data structure")

    def move(x):
        if isinstance(x, torch.Tensor):
            return x.to(map_location) if map_location is not None else x
        if isinstance(x, list):
            return [move(y) for y in x]
        if isinstance(x, DataStruct):
            x.data = move(x.data)
        return x

    if not isinstance(getattr(obj, "misc", None), defaultdict):
        obj.misc = defaultdict(_none, getattr(obj, "misc", None) or {})
    return move(obj)
