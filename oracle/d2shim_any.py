"""Permissive import stub -- TEST INFRASTRUCTURE (used by oracle/make_golden_step.py only, in this container).

`pt/engine/trainer.py` imports ~40 names from detectron2 / fvcore (trainers, hooks, evaluators, data loaders,
visualiser ...) that the step itself never touches. This finder fabricates any not-yet-registered
`detectron2.*` / `fvcore.*` (...) module on demand and answers every attribute with an inert dummy class, so that
the reference's trainer module can be IMPORTED unmodified; the methods on the hot path (`run_step`, `resize`,
`_update_teacher_model`, `clip_gradient`, `process_pseudo_label`, ...) are then executed for real on a namespace
object that carries the reference's own model classes (oracle/d2shim_model.py)."""
import importlib.abc
import importlib.machinery
import sys
import types


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Dummy

    def __call__(cls, *a, **k):
        if cls is Dummy:
            if len(a) == 1 and not k and callable(a[0]):
                return a[0]  # used as a decorator
            return type.__call__(cls)
        return type.__call__(cls, *a, **k)


class Dummy(metaclass=_Meta):
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Dummy

    def __call__(self, *a, **k):
        return Dummy()

    def __iter__(self):
        return iter(())


class AutoModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Dummy


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("detectron2", "fvcore", "pycocotools", "iopath", "yacs", "cv2", "tabulate", "termcolor")

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.ROOTS and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = AutoModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install():
    """Call AFTER oracle.d2shim(_model).install(): the real restatements stay, everything else becomes inert."""
    from probabilisticteacher_b200.structures import Boxes
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())
    for name, mod in list(sys.modules.items()):
        if name.split(".")[0] in _Finder.ROOTS and not isinstance(mod, AutoModule):
            mod.__class__ = AutoModule
            if not hasattr(mod, "__path__"):
                mod.__path__ = []
    boxes = AutoModule("detectron2.structures.boxes")
    boxes.__path__ = []
    boxes.Boxes = Boxes
    sys.modules["detectron2.structures.boxes"] = boxes
    comm = AutoModule("detectron2.utils.comm")
    comm.__path__ = []
    comm.get_world_size = lambda: 1
    comm.is_main_process = lambda: True
    comm.get_rank = lambda: 0
    sys.modules["detectron2.utils.comm"] = comm
    sys.modules["detectron2.utils"].comm = comm
