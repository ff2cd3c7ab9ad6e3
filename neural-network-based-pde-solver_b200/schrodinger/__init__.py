"""Drop-ins for the loss functions of the reference's Schrödinger scripts.

One module per reference script, same class / function names, argument order and return arity:

    ipw_1d_pinn_drm   Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_PINN_DRM.py
    ipw_1d_wan        Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN.py
    qho_2d            Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D.py
    kh_1d             Schrodinger_Equations/Kramers_Henneberger/KH_1D.py

The network classes keep the reference's layout (``.net`` Sequential, ``technique`` / ``enforce_bc`` /
``FN`` attributes), so the reference's own model objects are accepted as well.  Every loss runs in
the fused CUDA kernels (pde_b200.ops); nothing here differentiates through autograd graphs.
"""
from . import ipw_1d_pinn_drm, ipw_1d_wan, kh_1d, qho_2d

__all__ = ["ipw_1d_pinn_drm", "ipw_1d_wan", "qho_2d", "kh_1d"]
