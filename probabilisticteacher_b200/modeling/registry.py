"""Name -> builder registries mirroring the detectron2 registries the reference plugs into
(`pt/modeling/meta_arch/rcnn.py:30`, `backbone/vgg.py:189`, `proposal_generator/rpn.py:44,58`,
`anchor_generator.py:31`, `roi_heads/roi_heads.py:39`)."""


class Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o
        return deco(obj) if obj is not None else deco

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
RPN_HEAD_REGISTRY = Registry("RPN_HEAD")
ANCHOR_GENERATOR_REGISTRY = Registry("ANCHOR_GENERATOR")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
