"""Registries with the reference's decorator API (fairseq/registry.py:16-100, fairseq/models/__init__.py,
fairseq/tasks/__init__.py, fairseq/criterions/__init__.py): @register_model(name, dataclass=...),
@register_criterion, @register_task, @register_lr_scheduler, @register_audio_feature_transform.
Duplicate names raise ValueError like the reference.  When the real fairseq is importable, INTEGRATION.md
shows how the same classes are registered there through --user-dir.
"""
MODEL_REGISTRY, MODEL_DATACLASS_REGISTRY = {}, {}
CRITERION_REGISTRY, TASK_REGISTRY, LR_SCHEDULER_REGISTRY, FEATURE_TRANSFORM_REGISTRY = {}, {}, {}, {}


def _register(registry, kind):
    def deco_factory(name, dataclass=None):
        def deco(cls):
            if name in registry:
                raise ValueError("Cannot register duplicate %s (%s)" % (kind, name))
            registry[name] = cls
            if dataclass is not None:
                cls.__dataclass = dataclass
                if registry is MODEL_REGISTRY:
                    MODEL_DATACLASS_REGISTRY[name] = dataclass
            return cls
        return deco
    return deco_factory


register_model = _register(MODEL_REGISTRY, "model")
register_criterion = _register(CRITERION_REGISTRY, "criterion")
register_task = _register(TASK_REGISTRY, "task")
register_lr_scheduler = _register(LR_SCHEDULER_REGISTRY, "lr scheduler")
register_audio_feature_transform = _register(FEATURE_TRANSFORM_REGISTRY, "audio feature transform")
