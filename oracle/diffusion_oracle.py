"""Oracle (test infrastructure only): DDPM schedule, respacing and samplers of RoHM, restated on the CPU.

float64 numpy for every table (as the reference builds them), python ints for the respacing, torch-CPU fp32 for
the per-step tensor arithmetic.  Reference: diffusion/gaussian_diffusion_posenet.py, diffusion/
gaussian_diffusion_trajnet.py (identical arithmetic), diffusion/respace.py, utils/model_util.py.
"""
import math

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------------
# schedule (reference gaussian_diffusion_posenet.py:14-58)
# ----------------------------------------------------------------------------------------------------------
def named_beta_schedule(name, n, scale_betas=1.0):
    """Reference get_named_beta_schedule :14-38 and betas_for_alpha_bar :41-58."""
    if name == "linear":
        scale = scale_betas * 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if name == "cosine":
        def alpha_bar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        out = []
        for i in range(n):
            out.append(min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), 0.999))
        return np.array(out)
    raise NotImplementedError(name)


def diffusion_tables(betas):
    """All float64 coefficient tables of GaussianDiffusion*.__init__ (reference :131-168)."""
    betas = np.array(betas, dtype=np.float64)
    assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    ac_next = np.append(ac[1:], 0.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "alphas_cumprod_next": ac_next,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


# ----------------------------------------------------------------------------------------------------------
# respacing (reference respace.py:10-63, 76-90, 183-195)
# ----------------------------------------------------------------------------------------------------------
def space_timesteps(num_timesteps, section_counts):
    """Reference space_timesteps :10-63.  Returns a python set of ints (bit-exact requirement)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start = 0
    steps = []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))  # python round(): banker's rounding, as the reference
            cur += frac
        start += size
    return set(steps)


def respaced_betas(base_betas, use_timesteps):
    """SpacedDiffusion*.__init__ (reference respace.py:76-90): new betas + timestep_map."""
    use = set(use_timesteps)
    ac = np.cumprod(1.0 - np.array(base_betas, dtype=np.float64), axis=0)
    last = 1.0
    new_betas, tmap = [], []
    for i, a in enumerate(ac):
        if i in use:
            new_betas.append(1 - a / last)
            last = a
            tmap.append(i)
    return np.array(new_betas), tmap


def create_diffusion(noise_schedule, steps, timestep_respacing=""):
    """utils/model_util.py:6-40 reduced to what it computes: (tables, timestep_map)."""
    betas = named_beta_schedule(noise_schedule, steps, 1.0)
    if not timestep_respacing:
        timestep_respacing = [steps]
    nb, tmap = respaced_betas(betas, space_timesteps(steps, timestep_respacing))
    return diffusion_tables(nb), tmap


def extract(arr, i):
    """_extract_into_tensor (reference :967-980) for a batch-constant index: float64 table -> fp32 scalar."""
    return float(np.float32(arr[i]))


# ----------------------------------------------------------------------------------------------------------
# samplers.  model_fn(x_t, original_timestep:int) -> pred_xstart (torch fp32, same shape as x_t)
# ----------------------------------------------------------------------------------------------------------
def p_sample_step(tables, i, x_t, x0, noise, guidance=None):
    """One ancestral step (reference p_mean_variance :236-280 + p_sample :388-434 / p_sample_with_grad :436-480).

    guidance: optional list of (weight, grad) pairs; mean += weight * variance[i] * grad, applied in order.
    Every product is taken in fp32 exactly as the reference's broadcasting does.
    """
    c1 = extract(tables["posterior_mean_coef1"], i)
    c2 = extract(tables["posterior_mean_coef2"], i)
    var = extract(tables["posterior_variance"], i)
    logvar = extract(tables["posterior_log_variance_clipped"], i)
    mean = c1 * x0 + c2 * x_t
    if guidance:
        for w, g in guidance:
            mean = mean + w * torch.tensor(var, dtype=torch.float32) * g
    nonzero = 1.0 if i != 0 else 0.0
    sigma = torch.exp(0.5 * torch.tensor(logvar, dtype=torch.float32))
    return mean + nonzero * sigma * noise


def p_sample_loop(tables, tmap, model_fn, x_T, noise_fn, guidance_fn=None, early_stop=False):
    """p_sample_loop_progressive (reference :578-662).  noise_fn(i) -> fp32 noise tensor for step i (drawn every
    step incl. i == 0, as the reference does).  guidance_fn(i, x0) -> list of (weight, grad) or None.
    Returns (final sample or pred_xstart if early_stop, last x0)."""
    n = len(tables["betas"])
    indices = list(range(n))[::-1]
    if early_stop:
        indices = indices[0:980]
    x = x_T
    x0 = None
    for i in indices:
        x0 = model_fn(x, tmap[i])
        noise = noise_fn(i)
        g = guidance_fn(i, x0) if guidance_fn is not None else None
        x = p_sample_step(tables, i, x, x0, noise, g)
    return (x0 if early_stop else x), x0


def q_sample(tables, i, x_start, noise):
    """Reference q_sample :192-210."""
    return extract(tables["sqrt_alphas_cumprod"], i) * x_start + extract(tables["sqrt_one_minus_alphas_cumprod"], i) * noise


def ddim_step(tables, i, x_t, x0, noise, eta=0.0):
    """INTENDED DDIM update (reference ddim_sample :665-715, which cannot run as shipped -- SURVEY.md D4).
    Parity unpinned: restated from the math with `batch` threaded through."""
    f = lambda name: torch.tensor(extract(tables[name], i), dtype=torch.float32)
    eps = (f("sqrt_recip_alphas_cumprod") * x_t - x0) / f("sqrt_recipm1_alphas_cumprod")
    ab, ab_prev = f("alphas_cumprod"), f("alphas_cumprod_prev")
    sigma = eta * torch.sqrt((1 - ab_prev) / (1 - ab)) * torch.sqrt(1 - ab / ab_prev)
    mean = x0 * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - sigma ** 2) * eps
    nonzero = 1.0 if i != 0 else 0.0
    return mean + nonzero * sigma * noise


def ddim_sample_loop(tables, tmap, model_fn, x_T, noise_fn, eta=0.0):
    n = len(tables["betas"])
    x = x_T
    for i in list(range(n))[::-1]:
        x0 = model_fn(x, tmap[i])
        x = ddim_step(tables, i, x, x0, noise_fn(i), eta)
    return x
