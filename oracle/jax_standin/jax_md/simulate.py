import dataclasses as _dc



@_dc.dataclass
class BrownianState:
    position: object
    mass: object
    rng: object


def canonicalize_mass(state):
    return state
