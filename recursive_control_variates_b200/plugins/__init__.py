"""Plugin registration under the reference's names (plugins/__init__.py:11-17), plus the
aliases BASELINE.json's north_star uses (`cv_ps`, `twostatevolpath`)."""
from .. import mitsuba_compat as mi
from .cv_integrator import CVIntegrator
from .metaintegrator import MetaIntegrator
from .twostatemedium import TwoStateMedium
from .twostatevolpath import PRBVolpathIntegrator, TwoStatePRBVolpathIntegrator
from .volpathsimple import VolpathSimpleIntegrator
from .welford import StatisticsEstimator

mi.register_integrator("volpathsimple", lambda props: VolpathSimpleIntegrator(props))
mi.register_integrator("meta", MetaIntegrator)
mi.register_integrator("cv", CVIntegrator)
mi.register_integrator("twostateprbvolpath", TwoStatePRBVolpathIntegrator)
mi.register_integrator("prbvolpath", PRBVolpathIntegrator)
mi.register_medium("twostatemedium", TwoStateMedium)
# aliases
mi.register_integrator("twostatevolpath", TwoStatePRBVolpathIntegrator)
mi.register_integrator("cv_ps", lambda props: MetaIntegrator(mi.Properties({**props, "method": "cv_ps"})))

__all__ = ["CVIntegrator", "MetaIntegrator", "TwoStateMedium", "PRBVolpathIntegrator", "TwoStatePRBVolpathIntegrator",
           "VolpathSimpleIntegrator", "StatisticsEstimator"]
