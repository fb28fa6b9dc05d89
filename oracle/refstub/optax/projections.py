def projection_simplex(x, scale=1.0):
    raise NotImplementedError("refstub: optax.projections.projection_simplex is not on the BV MaxEnt path")
