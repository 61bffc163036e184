"""Conjugate gradient after Hestenes and Stiefel: examples/cghs.hh:15-45, with every scalar of the iteration resident where
the arrays live. The reference computes rho, sig, tau, t, gam as host doubles between array operations; here they are rank-0
views filled by reduce_into and combined by rank-0 maps, so an iteration is a chain of launches with ONE host read (the
convergence test the reference makes at the top of every iteration)."""
import numpy as np

from . import abi, ra


def cghs(mult, b, x, work, eps, scal, max_its=10 ** 9):
    """mult(v, w): w = A v.   b, x: arrays.   work: 3 x shape(b) scratch.   scal: a rank-1 view of 8 f64 slots.
    Returns the number of iterations, like the reference (examples/cghs.hh:44)."""
    work.assign(0.)
    g, r, p = work(0), work(1), work(2)
    err, gg, rho, sig, tau, t, gam = (scal(k) for k in range(7))
    ra.reduce_into(err, ra.sqrm(b))
    err *= float(eps) * float(eps)                      # err = sqr(eps)*reduce_sqrm(b)   (:26)
    mult(x, g)
    g.assign(b - g)
    r.assign(g)
    its = 0
    while its < max_its:
        ra.reduce_into(gg, ra.sqrm(g))
        e_gg = scal(ra.iota(2)).numpy()                 # the one host read per iteration: reduce_sqrm(g) > err   (:33)
        if not e_gg[1] > e_gg[0]:
            break
        mult(r, p)
        ra.reduce_into(rho, ra.sqrm(p))                 # :35-37
        ra.reduce_into(sig, r * p)
        ra.reduce_into(tau, g * r)
        t.assign(tau / sig)                             # :38
        gam.assign((t * t * rho - tau) / tau)           # :39
        x += t * r                                      # :40-42 (rank-0 operands repeat over every axis: prefix matching)
        g -= t * p
        r.assign(r * gam + g)
        its += 1
    return its
