"""Per-environment xi tables of the random-envs suite: the only env-specific data the DR samplers need.

For each env id: dimension, parameter names (``dyn_ind_to_name``), lowest feasible value per dim
(``get_task_lower_bound``) and ADR search bounds (``get_search_bounds_mean``).  The MuJoCo dynamics
themselves are out of scope; these tables let ``TaskSampler`` reproduce
``sample_task`` for every env from 3 to 30 dims.

Sources (reference tree): random_cartpole.py:104-147; jinja/random_hopper.py:42-72;
jinja/random_hopper_unmodeled.py:40-68; jinja/random_half_cheetah.py:46-83;
jinja/random_half_cheetah_unmodeled.py:43-80; jinja/random_walker2d.py:46-98;
jinja/random_walker2d_unmodeled.py:49-101; jinja/random_humanoid.py:55-148;
jinja/random_humanoid_unmodeled.py:70-164.
"""
from collections import namedtuple

XiTable = namedtuple("XiTable", "names lower_bounds search_bounds reward_threshold preferred_lr")

_MASS = (0.5, 10.0)


def _table(spec, reward_threshold, preferred_lr):
    names = tuple(n for n, _, _ in spec)
    return XiTable(names, tuple(float(lb) for _, lb, _ in spec), tuple((float(a), float(b)) for _, _, (a, b) in spec),
                   reward_threshold, preferred_lr)


def _masses(names, lb=0.1):
    return [(n, lb, _MASS) for n in names]


_cartpole = _table([("gravity", 0.1, (2.0, 20.0)), ("cart_mass", 0.1, (0.5, 3.0)),
                    ("pole_mass", 0.1, (0.05, 0.3)), ("pole_length", 0.1, (0.1, 1.0))], 500, None)

_hopper = _table(_masses(["torsomass", "thighmass", "legmass", "footmass"]), 1750, 0.0005)
_hopper_unmodeled = _table(_masses(["thighmass", "legmass", "footmass"], lb=0.001), 1750, 0.0005)

_friction = ("friction", 0.02, (0.1, 2.0))
_cheetah = _table(_masses(["torso", "bthigh", "bshin", "bfoot", "fthigh", "fshin", "ffoot"]) + [_friction], 4500, 0.0005)
_cheetah_unmodeled = _table(_masses(["bfoot", "fthigh", "fshin", "ffoot"]) + [_friction], 4500, 0.0005)

_walker_friction = [("friction_right", 0.05, (0.1, 3.0)), ("friction_left", 0.05, (0.1, 3.0))]
_walker = _table(_masses(["torso", "thigh", "leg", "foot", "thigh_left", "leg_left", "foot_left"])
                 + [(n, 0.1, (0.15, 1.0)) for n in ("torsosize", "thighsize", "legsize", "footsize")]
                 + _walker_friction, 2200, 0.0005)
_walker_unmodeled = _table(_masses(["foot", "thigh_left", "leg_left", "foot_left"])
                           + [("thighsize", 0.25, (0.3, 1.0)), ("legsize", 0.25, (0.3, 1.0)),
                              ("footsize", 0.12, (0.15, 0.8))]
                           + _walker_friction, 2200, 0.0005)

# humanoid dampers: joints 7 and 11..17 are the soft ones (knees / arms)
_SOFT = {7, 11, 12, 13, 14, 15, 16, 17}


def _damper(k):
    return ("damp%d" % k, 0.15, (0.2, 5.0)) if k in _SOFT else ("damp%d" % k, 0.8, (1.0, 10.0))


_humanoid = _table(_masses(["mass%d" % k for k in range(13)], lb=0.2) + [_damper(k) for k in range(1, 18)], 2200, 0.0001)
_humanoid_unmodeled = _table(_masses(["mass%d" % k for k in range(4, 13)], lb=0.2) + [_damper(k) for k in range(4, 18)],
                             2200, 0.0001)

XI_TABLES = {
    "RandomCartPole-v0": _cartpole,
    "RandomHopper-v0": _hopper, "RandomHopperNoisy-v0": _hopper,
    "RandomHopperUnmodeled-v0": _hopper_unmodeled,
    "RandomHalfCheetah-v0": _cheetah, "RandomHalfCheetahNoisy-v0": _cheetah,
    "RandomHalfCheetahUnmodeled-v0": _cheetah_unmodeled,
    "RandomWalker2d-v0": _walker, "RandomWalker2dNoisy-v0": _walker,
    "RandomWalker2dUnmodeled-v0": _walker_unmodeled,
    "RandomHumanoid-v0": _humanoid, "RandomHumanoidNoisy-v0": _humanoid,
    "RandomHumanoidUnmodeled-v0": _humanoid_unmodeled,
}

# Nominal humanoid xi for building plausible synthetic distributions (BASELINE config 5): the 17 joint
# dampings are in jinja/assets/humanoid.xml:38-87; the 13 body masses are computed by MuJoCo from geometry and
# are not in the reference tree (values of gym's Humanoid-v2 model, from memory) -- the sampler does not
# depend on them.
HUMANOID_NOMINAL = (8.322, 2.036, 5.853, 4.526, 2.632, 1.767, 4.526, 2.632, 1.767, 1.594, 1.198, 1.594, 1.198,
                    5.0, 5.0, 5.0, 5.0, 5.0, 5.0, 1.0, 5.0, 5.0, 5.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)


def get_table(env_id):
    try:
        return XI_TABLES[env_id]
    except KeyError:
        raise KeyError("no xi table for env id %r (known: %s)" % (env_id, ", ".join(sorted(XI_TABLES)))) from None
