"""AGV: the bicycle model with rear-wheel steering (dtheta = -V / L tan(delta)) and a
rectangular shape, so the heading enters the collision rows (reference
``omgtools/vehicles/agv.py``: bounds 50-63, trajectory constraints 73-109 -- the bicycle's
with the sign of every steering term flipped -- l'Hopital start constraint 111-144,
terminal constraints 146-167, parameters 200-224, collision constraints 226-240, ode
294-301).  Lowering as for the bicycle (vehicles/bicycle.py)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/agv.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .bicycle import Bicycle
from ..basics.shape import Rectangle


class AGV(Bicycle):

    def __init__(self, length=0.4, options=None, bounds=None, shapes=None):
        bounds = dict(bounds or {})
        bounds.setdefault('vmax', 0.5)
        Bicycle.__init__(self, length=length, options=options, bounds=bounds,
                         shapes=shapes if shapes is not None else Rectangle(width=0.8, height=0.2))
        self.steer_sign = -1.

    def set_default_options(self):
        Bicycle.set_default_options(self)
        self.options.update({'plot_type': 'agv'})
