class Tracer:  # nothing is ever traced here: isinstance(x, jax.core.Tracer) is always False
    pass
