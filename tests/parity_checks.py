"""Parity assertions shared by the CPU suite (gather oracle vs the reference's golden vectors) and
the GPU suite (CUDA path vs the same vectors, and vs the gather oracle at rounding level).
`make(tank_w, tank_h, h, capacity)` returns a backend with the sph_b200.Context call surface."""
import numpy as np

from common import (DENSITY_REL, ONE_STEP_MAX_H, ONE_STEP_RMS_H, ULPS_POS, load_golden, longrun_tolerances, pos_err_h,
                    ulp32, vel_err)


def fresh(make, name, warm):
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    b = make(tank_w, tank_h, h, len(st) + 64)
    b.set_params(t)
    b.upload(st)
    return b, z, t, tank_w, tank_h, h, st


def check_binning_and_neighbours_exact(make, name, warm):
    """cell ids, bucket order and neighbour sets: bit-exact (hash.c:35-47, :160-163, :185, :221)."""
    b, z, t, tank_w, tank_h, h, st = fresh(make, name, warm)
    uid, cells = b.cells()
    order = np.argsort(uid)
    assert np.array_equal(uid[order], np.arange(len(st)))
    assert np.array_equal(cells[order], z[f"w{warm}_cells"])
    # (bucket CONTENTS per reference cell follow from the per-uid cell ids; the device order itself is by
    #  sort-grid sub-cell, SPH_CELL_DIV per axis, then uid -- the order inside a reference bucket is not
    #  observable in a gather, only its contents are)
    assert np.array_equal(b.pairs(), z[f"w{warm}_pairs"])
    fu, fc = b.forward_counts()
    assert np.array_equal(fc[np.argsort(fu)], z[f"w{warm}_fwd"])


def check_stages_vs_reference(make, name, warm):
    """One step, stage by stage, from the reference's snapshot; tolerances in common.py."""
    b, z, t, tank_w, tank_h, h, st = fresh(make, name, warm)
    b.advect(); b.sort()
    a, _ = b.download()
    mx, rms = pos_err_h(a, z[f"w{warm}_advect"], h)
    assert mx <= ONE_STEP_MAX_H and rms <= ONE_STEP_RMS_H, ("advect", mx, rms)
    assert np.array_equal(a["x_prev"].view("u4"), st["x"].view("u4"))       # fluid.c:515-516
    b.density()
    a, _ = b.download()
    ref = z[f"w{warm}_density"]
    # same positions up to the viscosity-order difference: compare densities loosely here, tightly below
    assert np.abs(a["density"] - ref["density"]).max() <= 0.05 * max(1.0, ref["density"].max())
    b.relax(); b.sort()
    a, _ = b.download()
    mx, rms = pos_err_h(a, z[f"w{warm}_after1"], h)
    assert mx <= ONE_STEP_MAX_H and rms <= ONE_STEP_RMS_H, ("step", mx, rms)
    vmx, vrms = vel_err(a, z[f"w{warm}_after1"])
    assert vmx <= ONE_STEP_MAX_H * h / t.time_step and vrms <= ONE_STEP_RMS_H * h / t.time_step, ("vel", vmx, vrms)
    return mx, rms, vmx, vrms


def check_density_exact_positions(make, name, warm):
    """Density on the reference's OWN predicted positions (upload them): only the summation order
    differs from calculate_density (fluid.c:527-539), so rel 1e-5."""
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    ref = z[f"w{warm}_density"]
    b = make(tank_w, tank_h, h, len(ref) + 64)
    t0 = t.copy(); t0.g = 0.0; t0.sigma = 0.0; t0.beta = 0.0; t0.mover_type = bytes([2])   # advect becomes x += 0
    b.set_params(t0)
    st = ref.copy(); st["v_x"] = 0; st["v_y"] = 0
    b.upload(st)
    b.advect(); b.sort()
    # the lists hash_fluid(true) builds at the predicted positions (fluid.c:315): exact
    assert np.array_equal(b.pairs(), z[f"w{warm}_pairs_pred"])
    b.density()
    a, _ = b.download()
    assert np.array_equal(a["x"].view("u4"), ref["x"].view("u4")) and np.array_equal(a["y"].view("u4"), ref["y"].view("u4"))
    scale = max(1.0, float(ref["density"].max()))
    assert np.abs(a["density"] - ref["density"]).max() <= DENSITY_REL * scale
    assert np.abs(a["density_near"] - ref["density_near"]).max() <= DENSITY_REL * scale


def check_ten_steps_bounded(make, name, warm):
    """10 steps from the snapshot: chaotic growth, so only a loose bound."""
    b, z, t, tank_w, tank_h, h, st = fresh(make, name, warm)
    b.step(10)
    a, _ = b.download()
    mx, rms = pos_err_h(a, z[f"w{warm}_after10"], h)
    assert mx <= 0.5 and rms <= 2e-2, (mx, rms)
    return mx, rms


def density_of(make, tank_w, tank_h, h, t, aos):
    """Density of a state, recomputed from positions (same estimator for every implementation)."""
    b = make(tank_w, tank_h, h, len(aos) + 64)
    t0 = t.copy(); t0.g = 0.0; t0.sigma = 0.0; t0.beta = 0.0; t0.mover_type = bytes([2])
    b.set_params(t0)
    st = aos.copy(); st["v_x"] = 0; st["v_y"] = 0
    b.upload(st); b.advect(); b.sort(); b.density()
    return b.download()[0]["density"]


def check_long_run_statistics(make, name, lattice_state, dens_make=None, prepare=None, widen=None):
    """1200 steps from the lattice; statistics averaged over the last 200 agree with the reference's within
    common.longrun_tolerances(name) (the stated tolerances, widened to 1.5 x the reference's own sensitivity to
    its particle order where that is larger)."""
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    tol = longrun_tolerances(name)
    tol.update(widen or {})
    b = make(tank_w, tank_h, h, len(lattice_state) + 64)
    if prepare:
        prepare(b)
    b.set_params(t)
    b.upload(lattice_state)
    b.step(1000)
    acc = []
    for k in range(200):
        b.step(1)
        if k % 10 == 9:
            a, _ = b.download()
            d = density_of(dens_make or make, tank_w, tank_h, h, t, a)
            acc.append([d.mean(), d.max(), a["y"].mean(), 0.5 * (a["v_x"] ** 2 + a["v_y"] ** 2).mean()])
    got = np.array(acc).mean(axis=0)
    ref = z["longrun_stats"]
    assert abs(got[0] - ref[0]) <= tol["mean_density"] * ref[0], ("mean density", got, ref, tol)
    assert abs(got[1] - ref[1]) <= tol["max_density"] * ref[1], ("max density", got, ref, tol)
    assert abs(got[2] - ref[2]) <= tol["mean_height"] * ref[2], ("mean height", got, ref, tol)
    assert abs(got[3] - ref[3]) <= tol["ke_rel"] * ref[3] + tol["ke_abs"], ("KE", got, ref, tol)
    return got, ref


def check_tight_vs_gather_oracle(make, make_oracle, name, warm, steps=1):
    """CUDA vs the gather oracle: identical algorithm and summation order -> a few ulps."""
    b, z, t, tank_w, tank_h, h, st = fresh(make, name, warm)
    o, *_ = fresh(make_oracle, name, warm)
    tol = ULPS_POS * ulp32(tank_w)
    for s in range(steps):
        b.advect(); o.advect(); b.sort(); o.sort()
        a, ua = b.download(); r, ur = o.download()
        assert np.array_equal(ua, ur)
        assert np.abs(a["x"] - r["x"]).max() <= tol and np.abs(a["y"] - r["y"]).max() <= tol, ("advect", s)
        b.density(); o.density()
        a, _ = b.download(); r, _ = o.download()
        assert np.abs(a["density"] - r["density"]).max() <= DENSITY_REL * max(1.0, r["density"].max()) + 1e-4 * (s > 0)
        b.relax(); o.relax(); b.sort(); o.sort()
        a, _ = b.download(); r, _ = o.download()
        grow = 4 ** s      # rounding differences are amplified by the dynamics from step to step
        assert np.abs(a["x"] - r["x"]).max() <= tol * grow and np.abs(a["y"] - r["y"]).max() <= tol * grow, ("relax", s)
        assert np.abs(a["v_x"] - r["v_x"]).max() <= tol * grow / t.time_step * 1.5, ("vel", s)


def check_state_snapshot(make_ctx, as_sph):
    """sph_state_save / sph_state_restore: the steps after a restore are the steps after the save, bit for bit (graph
    and staged paths, a parameter change in between undone too)."""
    import sph_b200
    z, t, tank_w, tank_h, h, _ = load_golden("block3000")
    st = z["w150_state"]
    c = make_ctx(tank_w, tank_h, h, len(st) + 64)
    ts = as_sph(t)
    c.set_params(ts); c.upload(st); c.step(7)
    c.state_save()
    c.step(9)
    a, ua = c.download(order=sph_b200.ORDER_CELL)
    t2 = ts.copy(); t2.k = 0.5; t2.mover_center_x = 0.3 * tank_w
    c.set_params(t2); c.step(3)
    c.state_restore()
    c.step(5); c.advect(); c.sort(); c.density(); c.relax(); c.sort(); c.step(3)
    b, ub = c.download(order=sph_b200.ORDER_CELL)
    assert np.array_equal(ua, ub)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a[f].view("u4"), b[f].view("u4")), f


def check_clamped_impulses(make, make_oracle):
    """A violent goo state -- every second particle thrown at up to 5 units/s against its neighbours -- in which the +-5
    clamp of a pair's impulse (fluid.c:459) binds for many pairs: the trimmed candidate loop of k_advect must fall back
    to its exact body for those rows and agree with the gather oracle's clamped impulses."""
    z, t, tank_w, tank_h, h, _ = load_golden("goo_rect1508")
    st = z["w300_state"].copy()
    rng = np.random.default_rng(5)
    st["v_x"] = np.where(np.arange(len(st)) % 2 == 0, 5.0, -5.0).astype("f4") * rng.uniform(0.5, 1.0, len(st)).astype("f4")
    st["v_y"] = rng.uniform(-5, 5, len(st)).astype("f4")
    b = make(tank_w, tank_h, h, len(st) + 64); o = make_oracle(tank_w, tank_h, h, len(st) + 64)
    for s in (b, o):
        s.set_params(t); s.upload(st)
        s.advect(); s.sort()
    a, ua = b.download(); r, ur = o.download()
    assert np.array_equal(ua, ur)
    # predicted positions = x + v dt: a unit of velocity error is dt = 8.3e-3 of position
    assert np.abs(a["x"] - r["x"]).max() <= 1e-5 and np.abs(a["y"] - r["y"]).max() <= 1e-5
    x, y, vx, vy = (st[f].astype("f8") for f in ("x", "y", "v_x", "v_y"))
    dx = x[None, :] - x[:, None]; dy = y[None, :] - y[:, None]
    rr = np.hypot(dx, dy); np.fill_diagonal(rr, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = ((vx[:, None] - vx[None, :]) * dx + (vy[:, None] - vy[None, :]) * dy) / rr
        imp = 0.5 * t.time_step * (1 - rr / h) * (t.sigma * u + t.beta * u * u)
    binds = (rr <= h) & (u > 0) & ((np.abs(imp * dx / rr) > 2.5) | (np.abs(imp * dy / rr) > 2.5))
    assert binds.sum() > 20, binds.sum()
