"""Scene builders that talk to the mirrored `Solver`/`EntityFactory` API, so the same function feeds the CUDA path
and (in tests) the CPU oracle with identical f32 inputs.

Two kinds:
  * the reference's own live example scenes, restated call for call:
      setup_0_1_car_platformer  <- src/examples/0_1_car_platformer.zig + src/examples/utils.zig:6-54
      setup_0_3_many_boxes      <- src/examples/0_3_many_boxes.zig:12-61
    (all literal arithmetic is done in f32, as Zig does for `f32` operands)
  * the synthetic configurations of BASELINE.json / SURVEY.md §8(d): box1k, pile100k, mixed1M, pyramid20k and one
    world of batch4096x256.  PRNG = splitmix64, seed 0x5EED0000 + config index (+ world id), u = (x >> 40) * 2^-24.
"""
from __future__ import annotations

import numpy as np

from ._abi import SHAPE_DISC, SHAPE_RECT, body_desc_dtype
from .solver import BodyOptions, DiscOptions, Parameters, RectangleOptions

f32 = np.float32
DT = f32(1.0) / f32(60.0)


def _dist2(a, b):
    dx = f32(a[0]) - f32(b[0])
    dy = f32(a[1]) - f32(b[1])
    return np.sqrt(f32(dx * dx) + f32(dy * dy), dtype=f32)


# ---------------------------------------------------------------------------------------------------------------------
# reference example scenes
# ---------------------------------------------------------------------------------------------------------------------
def _car(fac, pos):
    """src/examples/utils.zig:6-54"""
    mu = f32(0.4)
    t_pos = (f32(pos[0]) - f32(5), f32(pos[1]) - f32(10))
    bh = fac.make_rectangle_body(BodyOptions(pos=pos, density=1, mu=mu), RectangleOptions(5, f32(0.9)))
    bh_pos = (f32(pos[0]), f32(pos[1]))
    mu = f32(1.5)
    rad = f32(1.0)
    wl_pos = (t_pos[0] + f32(3.5), t_pos[1] + (f32(9.8) - rad))
    wl = fac.make_disc_body(BodyOptions(pos=wl_pos, density=1, mu=mu), DiscOptions(rad))
    wr_pos = (t_pos[0] + f32(6.5), wl_pos[1])
    wr = fac.make_disc_body(BodyOptions(pos=wr_pos, density=1, mu=mu), DiscOptions(rad))
    mu = f32(0.5)
    dist_car = _dist2(bh_pos, wl_pos)
    bh2_pos = (t_pos[0] + f32(5.25), t_pos[1] + f32(10.8))
    bh2 = fac.make_rectangle_body(BodyOptions(pos=bh2_pos, density=1, mu=mu), RectangleOptions(f32(1.2), f32(0.5)))

    params = Parameters(beta=14, power_min=-2, power_max=2)
    fac.make_offset_distance_joint(params, wl, bh, (0, 0), (-1.5, 0), rad + f32(0.2))
    fac.make_offset_distance_joint(params, wr, bh, (0, 0), (1.5, 0), rad + f32(0.2))
    fac.make_distance_joint(params, wl, wr, 3)
    fac.make_distance_joint(params, wl, bh, dist_car)
    fac.make_distance_joint(params, wr, bh, dist_car)
    fac.exclude_collision_pair(bh, wl)
    fac.exclude_collision_pair(bh, wr)
    fac.exclude_collision_pair(bh, bh2)
    p1 = (bh_pos[0] + f32(0.25), bh_pos[1] + f32(0))
    dist21 = _dist2(p1, (bh2_pos[0] + f32(-1), bh2_pos[1] + f32(1)))
    dist22 = _dist2(p1, (bh2_pos[0] + f32(1), bh2_pos[1] + f32(1)))
    fac.make_offset_distance_joint(Parameters(), bh, bh2, (0.25, 0), (-1, 1), dist21)
    fac.make_offset_distance_joint(Parameters(), bh, bh2, (0.25, 0), (1, 1), dist22)
    dist23 = _dist2(bh_pos, bh2_pos)
    fac.make_distance_joint(Parameters(beta=100), bh, bh2, dist23)


def setup_0_1_car_platformer(solver):
    """src/examples/0_1_car_platformer.zig:12-259 (N = 111, 11 joints, 3 exclusion pairs)."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(f32(9.82))

    def rect(pos, w, h, mu, angle=0.0, static=False, density=1, mass=None):
        bo = BodyOptions(pos=pos, angle=f32(angle), mu=f32(mu), density=None if mass is not None else density,
                         mass=mass)
        hd = fac.make_rectangle_body(bo, RectangleOptions(f32(w), f32(h)))
        if static:
            hd.set_static(True)
        return hd

    def disc(pos, r, mu, static=False, density=1, mass=None):
        bo = BodyOptions(pos=pos, mu=f32(mu), density=None if mass is not None else density, mass=mass)
        hd = fac.make_disc_body(bo, DiscOptions(f32(r)))
        if static:
            hd.set_static(True)
        return hd

    rect((0, -100), 1000, 20, 0.3, static=True)
    rect((0, 0), 40, 10, 0.3, static=True)
    rect((-19, 15), 2, 20, 0.3, static=True)
    _car(fac, (5, 10))
    mu = 0.5
    for x in (-15, -14, -13, -12):
        for y in (9, 10, 11, 12, 13):
            rect((x, y), 0.6, 0.4, mu)
    rect((0, 5), 0.8, 0.4, mu, angle=1.0, static=True)
    rect((f32(4.1), 5), 0.8, 0.7, mu, angle=3.0, static=True)
    rect((1, 5), 0.8, 0.4, mu, angle=0.5, static=True)
    rect((15, 5), 0.3, 0.4, mu, angle=3.0, static=True)
    rect((-9, 5), 0.9, 0.3, mu, angle=-1.0, static=True)
    rect((-10, 5), 0.8, 0.6, mu, angle=-2.0, static=True)
    rect((20, 12), 15, 0.5, mu, static=True)
    rect((7, 14), 14, 0.5, mu, angle=-0.3, static=True)
    rect((30, 7), 20, 1.0, mu, angle=0.25, static=True)
    body = rect((35, 13), 10.5, 0.5, mu)
    fac.make_fixed_position_joint(Parameters(), body, (35, 13))
    disc((38, 19), 1.0, mu, mass=100)
    rect((48, 9.5), 15, 1, mu, static=True)
    rect((52, 12), 3, 3, mu, mass=10)
    rect((f32(76.2), f32(3.3)), 40, 1.0, mu, angle=-0.3, static=True)
    for xp in range(65, 71):
        for yp in range(9, 20):
            rect((xp, yp), 0.9, 0.4, 0.7)
    rect((110, -2), 35, 1.0, mu, static=True)
    disc((110, -4), 4.0, mu, static=True)
    body = rect((100, 5), 12.9, 0.5, mu)
    fac.make_fixed_position_joint(Parameters(), body, (100, 5))
    fac.make_motor_joint(Parameters(beta=100, power_max=100, power_min=-100), body, f32(3.14))
    disc((104, 8), 2.0, mu)


def drive_0_1(solver):
    """The key-handler inputs of the driven golden run (demos/native/src/main.zig:124-133), applied
    before every process() call."""
    solver.body_handle(4).set_ang_momentum(-20)
    solver.body_handle(5).set_ang_momentum(-20)
    solver.body_handle(3).set_torque(400)
    solver.body_handle(109).set_torque(50)
    st = solver.body_handle(3).get()
    solver.body_handle(3).set_force(3, st.force_y)


def setup_0_3_many_boxes(solver):
    """src/examples/0_3_many_boxes.zig:12-61 (N = 523, no joints)."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(f32(9.82))
    mu = f32(0.3)
    h = fac.make_rectangle_body(BodyOptions(pos=(20, -5), density=5, mu=mu), RectangleOptions(1000, 10))
    h.set_static(True)
    fac.make_rectangle_body(BodyOptions(pos=(-40, 10), vel=(70, 10), omega=-6, mass=100, mu=mu),
                            RectangleOptions(8.0, 8.0))
    fac.make_rectangle_body(BodyOptions(pos=(20, 5), density=5, mu=mu), RectangleOptions(4.0, 1.0))
    for x in range(10, 30):
        xf = f32(2.0) * f32(x)
        for y in range(10, 36):
            bo = BodyOptions(pos=(xf, f32(y)), density=5, mu=mu)
            if y % 2 == 0:
                fac.make_rectangle_body(bo, RectangleOptions(1.0, 1.0))
            else:
                fac.make_disc_body(bo, DiscOptions(0.5))


# ---------------------------------------------------------------------------------------------------------------------
# synthetic configurations
# ---------------------------------------------------------------------------------------------------------------------
class SplitMix64:
    """Vectorised splitmix64: draw k uniforms at once, bit-identical to the sequential generator."""
    GOLDEN = np.uint64(0x9E3779B97F4A7C15)

    def __init__(self, seed: int):
        self.state = np.uint64(seed)

    def next_u64(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            k = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + k * self.GOLDEN
            self.state = self.state + np.uint64(n) * self.GOLDEN
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.next_u64(n) >> np.uint64(40)).astype(np.float32) * f32(2.0 ** -24)
        return (f32(lo) + (f32(hi) - f32(lo)) * u).astype(np.float32)


GRAVITY, DENSITY, MU = f32(9.82), f32(5.0), f32(0.3)
CONFIG_SEED = {"box1k": 0x5EED0001, "pile100k": 0x5EED0002, "mixed1M": 0x5EED0003, "pyramid20k": 0x5EED0004,
               "batch4096x256": 0x5EED0005}


def _descs(n):
    d = np.zeros(n, body_desc_dtype())
    d["mu"] = MU
    d["mass_value"] = DENSITY
    d["mass_is_density"] = 1
    return d


def _static_rect(pos, w, h):
    d = _descs(1)
    d["pos_x"], d["pos_y"] = pos
    d["shape"], d["a"], d["b"], d["is_static"] = SHAPE_RECT, w, h, 1
    return d


def _lattice(nx, ny, pitch, origin, scale=1.0):
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")  # row-major: iy outer, ix inner
    ix, iy = ix.ravel(), iy.ravel()
    x = (f32(origin[0]) + ix.astype(np.float32) * f32(pitch * scale)).astype(np.float32)
    y = (f32(origin[1]) + iy.astype(np.float32) * f32(pitch * scale)).astype(np.float32)
    return ix, iy, x, y


def descs_box(rng, nx, ny, pitch=1.25, origin=(-24.375, 2.0)):
    """nx*ny mixed discs (r 0.5) / rects (1x1) on a lattice, cfg1 style (mix mirrors 0_3_many_boxes.zig:45-60)."""
    ix, iy, x, y = _lattice(nx, ny, pitch, origin)
    n = nx * ny
    d = _descs(n)
    jit = rng.uniform(3 * n, -1.0, 1.0).reshape(n, 3)
    d["pos_x"] = x + jit[:, 0] * f32(0.05)
    d["pos_y"] = y + jit[:, 1] * f32(0.05)
    d["angle"] = jit[:, 2] * f32(0.1)
    is_disc = ((ix + iy) % 2) == 0
    d["shape"] = np.where(is_disc, SHAPE_DISC, SHAPE_RECT)
    d["a"] = np.where(is_disc, f32(0.5), f32(1.0))
    d["b"] = np.where(is_disc, f32(0.0), f32(1.0))
    return d


def build_box1k(solver, seed=CONFIG_SEED["box1k"]):
    """cfg1: 1,000 mixed bodies falling into a static box (3 static + 1000 dynamic); S=4, I=4."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(GRAVITY)
    rng = SplitMix64(seed)
    statics = np.concatenate([_static_rect((0, -1), 60, 2), _static_rect((-31, 30), 2, 60), _static_rect((31, 30), 2, 60)])
    d = np.concatenate([statics, descs_box(rng, 40, 25)])
    fac.make_bodies(d)
    return {"sub_steps": 4, "iters": 4, "n_rect": int(np.count_nonzero(d["shape"] == SHAPE_RECT))}


def build_pile(solver, nx=400, ny=250, seed=CONFIG_SEED["pile100k"]):
    """cfg2: nx*ny discs r in U(0.3, 0.5) on a pitch-1.1 lattice inside a floor and two walls (pile100k = 400x250)."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(GRAVITY)
    rng = SplitMix64(seed)
    pitch = 1.1
    half = nx * pitch / 2.0
    statics = np.concatenate([
        _static_rect((0, -1), 2 * half + 4, 2),
        _static_rect((-(half + 2), ny * pitch * 0.6), 2, ny * pitch * 1.2 + 4),
        _static_rect((half + 2, ny * pitch * 0.6), 2, ny * pitch * 1.2 + 4)])
    ix, iy, x, y = _lattice(nx, ny, pitch, (-(nx - 1) * pitch / 2.0, 0.6))
    n = nx * ny
    d = _descs(n)
    r = rng.uniform(3 * n).reshape(n, 3)
    d["pos_x"] = x + (r[:, 0] - f32(0.5)) * f32(0.1)
    d["pos_y"] = y + (r[:, 1] - f32(0.5)) * f32(0.1)
    d["shape"] = SHAPE_DISC
    d["a"] = f32(0.3) + r[:, 2] * f32(0.2)
    fac.make_bodies(np.concatenate([statics, d]))
    return {"sub_steps": 4, "iters": 4, "n_rect": 3}


def build_pile100k(solver):
    """cfg2.  2000 columns x 50 rows: a wide, shallow drop so that the pile has actually formed after ~200 process()
    calls (SURVEY 8d's 400 x 250 lattice is 275 m tall and is still in free fall at step 100 — measured in
    profiles/r01_pile_evolution.md); same body count, radii, pitch, jitter and materials."""
    return build_pile(solver, 2000, 50)


def build_mixed(solver, nx=1250, ny=800, n_large=1000, seed=CONFIG_SEED["mixed1M"]):
    """cfg3: nx*ny bodies, half discs r in U(0.25,0.5), half rects w,h in U(0.4,0.7) at any angle, plus `n_large`
    rects of 8-16 m on a coarse pitch-20 lattice above the field (multi-cell grid entries)."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(GRAVITY)
    rng = SplitMix64(seed)
    pitch = 1.1
    half = nx * pitch / 2.0
    top = ny * pitch
    statics = np.concatenate([
        _static_rect((0, -1), 2 * half + 4, 2),
        _static_rect((-(half + 2), top * 0.75), 2, top * 1.5 + 4),
        _static_rect((half + 2, top * 0.75), 2, top * 1.5 + 4)])
    ix, iy, x, y = _lattice(nx, ny, pitch, (-(nx - 1) * pitch / 2.0, 0.6))
    n = nx * ny
    d = _descs(n)
    r = rng.uniform(5 * n).reshape(n, 5)
    d["pos_x"], d["pos_y"] = x, y
    is_disc = r[:, 0] < f32(0.5)
    d["shape"] = np.where(is_disc, SHAPE_DISC, SHAPE_RECT)
    d["a"] = np.where(is_disc, f32(0.25) + r[:, 1] * f32(0.25), f32(0.4) + r[:, 1] * f32(0.3))
    d["b"] = np.where(is_disc, f32(0.0), f32(0.4) + r[:, 2] * f32(0.3))
    d["angle"] = np.where(is_disc, f32(0.0), (r[:, 3] * f32(2.0) - f32(1.0)) * f32(np.pi))
    parts = [statics, d]
    if n_large:
        lx = max(1, min(n_large, int((2 * half - 40) // 20)))
        ly = (n_large + lx - 1) // lx
        _, _, bx, by = _lattice(lx, ly, 20.0, (-(lx - 1) * 10.0, top + 20.0))
        big = _descs(lx * ly)[:n_large]
        rb = rng.uniform(3 * n_large).reshape(n_large, 3)
        big["pos_x"], big["pos_y"] = bx[:n_large], by[:n_large]
        big["shape"] = SHAPE_RECT
        big["a"] = f32(8.0) + rb[:, 0] * f32(8.0)
        big["b"] = f32(8.0) + rb[:, 1] * f32(8.0)
        big["angle"] = (rb[:, 2] * f32(2.0) - f32(1.0)) * f32(np.pi)
        parts.append(big)
    allb = np.concatenate(parts)
    fac.make_bodies(allb)
    return {"sub_steps": 4, "iters": 4, "n_rect": int(np.count_nonzero(allb["shape"] == SHAPE_RECT))}


def build_mixed1M(solver):
    """cfg3.  20000 x 50 lattice + 1000 large rectangles above it: as wide and shallow as pile100k (2000 x 50), so that a
    pile HAS formed after the 200-call pre-roll (SURVEY 8d's 1250 x 800 lattice is 880 m tall and round 1's 5000 x 200 is
    220 m: both are still in free fall when they are timed — profiles/r02_mixed1M_evolution.md); same body count, size
    distributions, pitch and materials."""
    return build_mixed(solver, 20000, 50, 1000)


def build_hub(solver, n_discs=48):
    """A plank resting on a row of `n_discs` discs: ONE non-static body with n_discs manifolds (as many colours), the
    case the per-body manifold lists of the dataflow colouring cannot hold (they fall back to colouring by rounds)."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(GRAVITY)
    pitch, r = 0.85, 0.4
    width = n_discs * pitch
    d = _descs(n_discs)
    d["pos_x"] = (np.arange(n_discs, dtype=np.float32) - f32((n_discs - 1) / 2.0)) * f32(pitch)
    d["pos_y"] = f32(r)
    d["shape"], d["a"] = SHAPE_DISC, f32(r)
    plank = _descs(1)
    plank["pos_x"], plank["pos_y"] = 0.0, 2 * r + 0.3
    plank["shape"], plank["a"], plank["b"] = SHAPE_RECT, width, 0.6
    fac.make_bodies(np.concatenate([_static_rect((0, -1), width + 8, 2), d, plank]))
    return {"sub_steps": 4, "iters": 4}


def build_pyramid(solver, base=199, n_spinners=50, seed=CONFIG_SEED["pyramid20k"]):
    """cfg4: rectangle pyramid (1.0 x 0.5, base row `base`), distance joints between horizontal neighbours on every
    10th row, fixed-position joints on both ends of the base row, `n_spinners` motor-driven bars beside it
    (joint parameters from src/examples/0_1_car_platformer.zig:145-148,245-255); S=4, I=10."""
    fac = solver.entity_factory()
    fac.make_downwards_gravity(GRAVITY)
    width = base * 1.05 + 40 + n_spinners * 12.0
    floor = _static_rect((0, -1), 2 * width, 2)
    rows = []
    row_start = []
    count = 0
    for r in range(base):
        nrow = base - r
        xs = (np.arange(nrow, dtype=np.float32) - f32((nrow - 1) / 2.0)) * f32(1.05)
        d = _descs(nrow)
        d["pos_x"] = xs
        d["pos_y"] = f32(0.25) + f32(r) * f32(0.5)
        d["shape"], d["a"], d["b"] = SHAPE_RECT, 1.0, 0.5
        rows.append(d)
        row_start.append(count)
        count += nrow
    first = fac.make_bodies(np.concatenate([floor] + rows))
    body0 = first + 1
    for r in range(0, base, 10):
        nrow = base - r
        for k in range(nrow - 1):
            a = body0 + row_start[r] + k
            fac.make_distance_joint(Parameters(), solver.body_handle(a), solver.body_handle(a + 1), f32(1.05))
    nrow = base
    for k in (0, nrow - 1):
        st = rows[0][k]
        fac.make_fixed_position_joint(Parameters(), solver.body_handle(body0 + k), (st["pos_x"], st["pos_y"]))
    x0 = base * 1.05 / 2.0 + 12.0
    for k in range(n_spinners):
        pos = (f32(x0 + 12.0 * k), f32(6.0))
        h = fac.make_rectangle_body(BodyOptions(pos=pos, density=DENSITY, mu=MU), RectangleOptions(10.5, 0.5))
        fac.make_fixed_position_joint(Parameters(), h, pos)
        fac.make_motor_joint(Parameters(beta=100, power_max=100, power_min=-100), h, f32(3.14))
    return {"sub_steps": 4, "iters": 10, "n_rect": count + 1 + n_spinners}


def build_pyramid20k(solver):
    return build_pyramid(solver, 199, 50)


def descs_batch_world(world_id: int, nx=23, ny=11, seed=CONFIG_SEED["batch4096x256"]):
    """One world of cfg5: 3 static box rects + nx*ny (= 253) dynamic mixed bodies, per-world seed."""
    rng = SplitMix64(seed + world_id)
    statics = np.concatenate([_static_rect((0, -1), 36, 2), _static_rect((-17, 15), 2, 30), _static_rect((17, 15), 2, 30)])
    return np.concatenate([statics, descs_box(rng, nx, ny, pitch=1.25, origin=(-13.75, 2.0))])


def build_batch_world(solver, world_id: int, nx=23, ny=11):
    fac = solver.entity_factory()
    fac.make_downwards_gravity(GRAVITY)
    fac.make_bodies(descs_batch_world(world_id, nx, ny))
    return {"sub_steps": 4, "iters": 4}
