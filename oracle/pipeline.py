"""ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's generate loop (csrc/libsdod/src/context.cpp:292-403) restated
over the fp32 torch networks (ldm_oracle) and the C sampler restatement (sampler_oracle.c): per step two UNet
evaluations (cond, uncond: context.cpp:352,366), CFG combine (:362,373), DPMSolver::update (:378); then the decoder
and the uint8 map (:386-395).  Latent NCHW [n,4,S,S]."""
import numpy as np
import torch

from . import sampler as S


@torch.no_grad()
def generate(unet, vae, cond, uncond, latents, guidance=7.5, steps=20, device="cpu", return_trace=False, sampler="dpm"):
    solver = S.OracleSolver()
    solver.prepare(steps)
    model_ts = solver.table("model_ts")[:steps]
    if sampler in ("ddim", "plms"):                                       # row f4; public CompVis ddim.py / plms.py, parity unpinned
        ddim_t, ddim_a, ddim_ap = S.ddim_tables(steps)
        model_ts = ddim_t.astype(np.float32)
    n = latents.shape[0]
    x = latents.clone().float().cpu().numpy()
    unet, vae = unet.to(device), vae.to(device)
    cond, uncond = cond.to(device).float(), (None if uncond is None else uncond.to(device).float())
    emb_all = unet.embed_time(torch.tensor(model_ts, device=device))
    solvers = [S.OracleSolver() for _ in range(n)]                       # one libsdod context per image
    for s in solvers:
        s.prepare(steps)
    trace = []

    def model_eps(xn, step):
        xt = torch.from_numpy(xn).to(device)
        emb = emb_all[step:step + 1].expand(n, -1)
        e_c = unet(xt, emb, cond).cpu().numpy()
        if guidance == 1.0:
            return e_c
        e_u = unet(xt, emb, uncond).cpu().numpy()
        return np.stack([S.cfg_combine(e_c[i].ravel(), e_u[i].ravel(), guidance).reshape(e_c[i].shape) for i in range(n)]).astype(np.float32)

    if sampler == "plms":                                                  # plms.py p_sample_plms
        old_eps = []
        for step in range(steps):
            e_t = model_eps(x, step)
            if not old_eps and steps > 1:
                x_prev = S.ddim_update(x, e_t, ddim_a[step], ddim_ap[step])
                e_next = model_eps(x_prev, step + 1)
                e_prime = np.float32(0.5) * e_t + np.float32(0.5) * e_next
            elif not old_eps:
                e_prime = e_t
            else:
                w = S.plms_step_weights(len(old_eps))
                e_prime = np.float32(w[0]) * e_t
                for wk, h in zip(w[1:], reversed(old_eps)):
                    e_prime = e_prime + np.float32(wk) * h
            x = S.ddim_update(x, e_prime.astype(np.float32), ddim_a[step], ddim_ap[step])
            old_eps = (old_eps + [e_t])[-3:]
            if return_trace:
                trace.append(x.copy())
        steps = 0                                                          # skip the DPM / DDIM loop below
    for step in range(steps):
        xt = torch.from_numpy(x).to(device)
        emb = emb_all[step:step + 1].expand(n, -1)
        e_c = unet(xt, emb, cond).cpu().numpy()
        if guidance == 1.0:
            e = e_c
        else:
            e_u = unet(xt, emb, uncond).cpu().numpy()
            e = np.stack([S.cfg_combine(e_c[i].ravel(), e_u[i].ravel(), guidance).reshape(e_c[i].shape) for i in range(n)])
        for i in range(n):
            if sampler == "ddim":
                x[i] = S.ddim_update(x[i], e[i].astype(np.float32), ddim_a[step], ddim_ap[step])
                continue
            xi, ei = np.ascontiguousarray(x[i]).ravel(), np.ascontiguousarray(e[i]).ravel().copy()
            solvers[i].update(step, xi, ei)
            x[i] = xi.reshape(x[i].shape)
        if return_trace:
            trace.append(x.copy())
    img = vae(torch.from_numpy(x).to(device)).permute(0, 2, 3, 1).cpu().numpy()       # [n,H,W,3] in [0,1]
    u8 = S.to_u8(img)
    return (u8, img, x, trace) if return_trace else (u8, img, x)
