"""Drop-ins for the loss functions of the reference's Schrödinger scripts.

One module per reference script, same class / function names, argument order and return arity:

    ipw_1d_pinn_drm   Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_PINN_DRM.py
    ipw_1d_wan        Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN.py
    qho_2d            Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D.py
    kh_1d             Schrodinger_Equations/Kramers_Henneberger/KH_1D.py
    ipw_1d_wan_fn     Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN_FN.py
    ipw_2d            Schrodinger_Equations/Infinite_Potential_Well/IPW_2D.py
    qho_1d_pinn_drm   Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_1D_PINN_DRM.py
    qho_1d_wan        Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_1D_WAN.py
    qho_2d_energy     Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D_Energy.py

The network classes keep the reference's layout (``.net`` Sequential, ``technique`` / ``enforce_bc`` /
``FN`` attributes), so the reference's own model objects are accepted as well.  Every loss runs in
the fused CUDA kernels (pde_b200.ops); nothing here differentiates through autograd graphs.
"""
from . import ipw_1d_pinn_drm, ipw_1d_wan, ipw_1d_wan_fn, ipw_2d, kh_1d, qho_1d_pinn_drm, qho_1d_wan, qho_2d, qho_2d_energy

__all__ = ["ipw_1d_pinn_drm", "ipw_1d_wan", "ipw_1d_wan_fn", "ipw_2d", "qho_1d_pinn_drm", "qho_1d_wan", "qho_2d",
           "qho_2d_energy", "kh_1d"]
