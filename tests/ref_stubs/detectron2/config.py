def configurable(init_func=None, *, from_config=None):
    """detectron2.config.configurable without the config machinery: explicit kwargs only."""
    if init_func is not None:
        return init_func
    return lambda f: f
