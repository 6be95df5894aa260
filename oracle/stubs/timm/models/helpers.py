def adapt_input_conv(in_chans, conv_weight):
    return conv_weight


def build_model_with_cfg(model_cls, variant, pretrained, default_cfg=None, **kwargs):
    kwargs.pop("pretrained_filter_fn", None)
    kwargs.pop("pretrained_custom_load", None)
    kwargs.pop("representation_size", None) if kwargs.get("representation_size", 1) is None else None
    model = model_cls(**kwargs)
    model.default_cfg = default_cfg
    return model


def named_apply(fn, module, name="", depth_first=True, include_root=False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = ".".join((name, child_name)) if name else child_name
        named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module
