"""Host-side diffusion schedule: float64 coefficient tables, timestep respacing, per-step coefficient rows.

Pure numpy / python ints (init-time only; reference: diffusion/gaussian_diffusion_posenet.py:14-58, 114-168,
diffusion/respace.py:10-63, 76-90).  The float64 tables are the contract: every value handed to a kernel is the
float64 entry rounded once to fp32, exactly what the reference's ``_extract_into_tensor(...).float()`` does (:977).
"""
import math

import numpy as np


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.0):
    """'linear' (Ho et al.) or 'cosine' (Nichol & Dhariwal) beta schedule as float64 [N]."""
    n = int(num_diffusion_timesteps)
    if schedule_name == "linear":
        scale = scale_betas * 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(n, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    n = int(num_diffusion_timesteps)
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


def space_timesteps(num_timesteps, section_counts):
    """Set of retained original timesteps.  "ddimN": first integer stride giving exactly N steps (ValueError if none);
    list / comma string: per-section fractional stride with python round() (banker's rounding)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    n_sections = len(section_counts)
    size_per, extra = divmod(num_timesteps, n_sections)
    start, taken = 0, []
    for idx, count in enumerate(section_counts):
        size = size_per + (1 if idx < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            taken.append(start + round(pos))
            pos += stride
        start += size
    return set(taken)


def respace(betas, use_timesteps):
    """(new_betas float64, timestep_map list[int]) of the process that visits only `use_timesteps`."""
    keep = set(use_timesteps)
    alphas_cumprod = np.cumprod(1.0 - np.asarray(betas, dtype=np.float64), axis=0)
    prev, new_betas, tmap = 1.0, [], []
    for i, ac in enumerate(alphas_cumprod):
        if i in keep:
            new_betas.append(1 - ac / prev)
            prev = ac
            tmap.append(i)
    return np.array(new_betas), tmap


TABLE_NAMES = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2")


def build_tables(betas):
    """dict of the 13 float64 tables (attribute names of the reference diffusion object)."""
    betas = np.array(betas, dtype=np.float64)
    assert len(betas.shape) == 1, "betas must be 1-D"
    assert (betas > 0).all() and (betas <= 1).all()
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    ac_next = np.append(ac[1:], 0.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    t = {
        "betas": betas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev, "alphas_cumprod_next": ac_next,
        "sqrt_alphas_cumprod": np.sqrt(ac), "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac), "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1), "posterior_variance": post_var,
        # log is clipped because the posterior variance is 0 at the start of the chain
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    assert ac_prev.shape == (len(betas),)
    return t


def ddpm_coef_rows(tables):
    """fp32 [N, 8] rows {coef1, coef2, sigma, variance, 0, 0, 0, 0} for rohm_ddpm_step, one per step index.
    sigma = (i != 0) * exp(0.5 * fp32(logvar)) evaluated in fp32 like the reference (:433)."""
    n = len(tables["betas"])
    rows = np.zeros((n, 8), dtype=np.float32)
    rows[:, 0] = tables["posterior_mean_coef1"].astype(np.float32)
    rows[:, 1] = tables["posterior_mean_coef2"].astype(np.float32)
    import torch
    logvar = torch.from_numpy(tables["posterior_log_variance_clipped"]).float()
    sigma = torch.exp(0.5 * logvar).numpy().copy()  # torch's fp32 exp, the function the reference calls
    sigma[0] = 0.0
    rows[:, 2] = sigma
    rows[:, 3] = tables["posterior_variance"].astype(np.float32)
    return rows


def ddim_coefs(tables, i, eta=0.0):
    """fp32 scalars of the intended DDIM update (see rohm_ddim_step), computed in fp32 like torch would."""
    f = lambda name: np.float32(tables[name][i])
    ab, ab_prev = f("alphas_cumprod"), f("alphas_cumprod_prev")
    sigma = np.float32(eta) * np.sqrt((1 - ab_prev) / (1 - ab)) * np.sqrt(1 - ab / ab_prev)
    dir_coef = np.sqrt(np.float32(1) - ab_prev - sigma * sigma)
    return (float(f("sqrt_recip_alphas_cumprod")), float(f("sqrt_recipm1_alphas_cumprod")), float(np.sqrt(ab_prev)),
            float(dir_coef), float(sigma if i != 0 else np.float32(0)))
