_REGISTRY = {}


def register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


def create_model(name, pretrained=False, **kwargs):
    return _REGISTRY[name](pretrained=pretrained, **kwargs)
