"""The reference's plugin API: ONE `Registry('models')` aliased for every kind of module
(mmdet/models/builder.py:7-15), classes self-register with `@X.register_module()`, and
`build_detector(cfg.model, train_cfg, test_cfg)` (builder.py:48-59, called from tools/DA_train.py:265-269)
instantiates `cfg['type']` with the remaining keys as constructor kwargs.  mmcv is not importable here, so
this is a minimal, behaviour-compatible stand-in for mmcv.utils.Registry on the DA path."""
import inspect
import warnings


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f"{self.__class__.__name__}(name={self._name}, items={sorted(self._module_dict)})"

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        names = [name or cls.__name__] if not isinstance(name, (list, tuple)) else list(name)
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f"{n} is already registered in {self._name}")
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _wrap(cls):
            self._register(cls, name, force)
            return cls

        return _wrap

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError(f"`cfg` must be a dict containing the key 'type', got {cfg!r}")
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop("type")
        if isinstance(obj_type, str):
            cls = self.get(obj_type)
            if cls is None:
                raise KeyError(f"{obj_type} is not in the {self._name} registry")
        elif inspect.isclass(obj_type):
            cls = obj_type
        else:
            raise TypeError(f"type must be a str or class, got {type(obj_type)}")
        try:
            return cls(**args)
        except Exception as e:
            raise type(e)(f"{cls.__name__}: {e}")


MODELS = Registry("models")
BACKBONES = NECKS = ROI_EXTRACTORS = SHARED_HEADS = HEADS = LOSSES = DETECTORS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_roi_extractor(cfg):
    return ROI_EXTRACTORS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """mmdet/models/builder.py:48-59."""
    if train_cfg is not None or test_cfg is not None:
        warnings.warn("train_cfg and test_cfg is deprecated, please specify them in model", UserWarning)
    assert cfg.get("train_cfg") is None or train_cfg is None, "train_cfg specified in both outer field and model field"
    assert cfg.get("test_cfg") is None or test_cfg is None, "test_cfg specified in both outer field and model field"
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))
