"""Named smoothers (sgmcmc_ssm/particle_filters/pf.py).  The per-step numpy functions of the reference
are fused into the CUDA step kernels; these objects only carry the `pf` name so that
`pf_wrapper(smoother=nemeth_smoother, ...)` keeps working."""


def _named(pf_name, doc):
    def smoother(*args, **kwargs):
        raise NotImplementedError(
            "the per-step smoother runs inside the CUDA library; call buffered_pf_wrapper(pf='%s', ...)" % pf_name)
    smoother.pf_name = pf_name
    smoother.__doc__ = doc
    smoother.__name__ = pf_name
    return smoother


nemeth_smoother = _named("nemeth", "Nemeth et al. O(N) (pf.py:138-181); Poyiadjis O(N) when lambduh=1")
poyiadjis_smoother = _named("poyiadjis_N2", "Poyiadjis et al. O(N^2) (pf.py:84-136)")
paris_smoother = _named("paris", "PaRIS (pf.py:183-341)")
pf_filter = _named("filter", "bootstrap filter statistic (pf.py:40-82)")
