"""gprot-fit: the reference's command line (scripts/gprot-fit:119-213) on the B200 engine.

Same flags and meaning.  What changes underneath: ``--ncores`` / ``--mpi`` no longer fan walkers out to worker
processes -- the ensemble is evaluated as batched launches on the GPU of this process (``--device``).  Stars are
fitted one after another and a failure of one star is logged and skipped, as in the reference (:199-213).
Data sources: ``--aigrain`` reads the Aigrain et al. simulation files (AIGRAIN_ROTATION, gprot/config.py);
``--kepler`` needs kplr / network access and the bandpass filter and ACF prior need the out-of-scope ACF code, so
those options exit with a clear message instead of silently doing something else.  MultiNest / PolyChord are not
installed here (the reference's polychord path raises NotImplementedError as well, scripts/gprot-fit:14-15).
"""
from __future__ import annotations

import argparse
import logging
import os
import sys

import numpy as np


def get_model(i, aigrain=True, kepler=False, ndays=None, subsample=40, chunksize=200, daterange=None,
              resultsdir='results', quarters=None, clever=True, bestchunk=None, pmax=None, filter=False, tag=None,
              altmodel=False, nochunks=False, npoints=600, acf_prior=False, sap=False, offline=False, device=0,
              **kwargs):
    """scripts/gprot-fit:32-98 (the light-curve and prior-diagnostic plots need matplotlib and are skipped)."""
    from .lc import AigrainLightCurve
    from .model import GPRotModel

    if not aigrain and not kepler:
        raise ValueError('Must specify either --aigrain or --kepler.')
    if kepler:
        raise NotImplementedError('--kepler needs kplr and network access to download light curves (gprot/kepler.py); '
                                  'build a LightCurve from your own arrays and use gprotation_b200.fit.fit_emcee3.')
    if filter or acf_prior:
        raise NotImplementedError('--filter / --acf_prior need the bandpass-filter + ACF pre-processing, which is outside '
                                  'the likelihood path (SURVEY.md section 2); pass period_mixture= to GPRotModel instead.')
    if altmodel:
        raise NotImplementedError('--altmodel (GPRotModel2, cosine kernel) is not part of the accelerated path.')
    if quarters is not None:
        raise NotImplementedError('--quarters needs the Kepler quarter table (gprot/data/qStartStop.txt).')
    if clever:
        daterange = None
        subsample = None
    if nochunks:
        chunksize = None
    lc = AigrainLightCurve(i, ndays, subsample, chunksize=chunksize)
    if clever and bestchunk is None:
        lc.make_best_chunks(npoints=npoints, chunksize=chunksize)
    if bestchunk is not None:
        lc.make_best_chunks(bestchunk)
    if daterange is not None:
        lc.restrict_range(daterange)
    if tag is not None:
        lc.name = lc.name + '_{}'.format(tag)
    if pmax is None and bestchunk is not None:
        try:
            pmax = np.log(max(bestchunk))
        except TypeError:
            pmax = np.log(bestchunk)
    mod = GPRotModel(lc, pmax=pmax, acf_prior=acf_prior, device=device)
    if not os.path.exists(resultsdir):
        os.makedirs(resultsdir)
    return mod


def _fit_emcee3(i, **kwargs):
    from .fit import fit_emcee3

    mod = get_model(i, **kwargs)
    return fit_emcee3(mod, **kwargs)


def build_parser():
    parser = argparse.ArgumentParser(description='Quasi-periodic GP rotation-period fits of Aigrain simulations or Kepler light curves on the B200 engine.')
    parser.add_argument('stars', nargs='+', type=int, help='Star ids, fitted one after another.')
    datagroup = parser.add_mutually_exclusive_group()
    datagroup.add_argument("--aigrain", dest="aigrain", action='store_true', help="Read the Aigrain et al. simulation files (AIGRAIN_ROTATION).")
    datagroup.add_argument("--kepler", dest="kepler", action='store_true', help="Kepler light curves (needs kplr; not available offline).")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--ncores", dest="n_cores", default=1, type=int,
                       help="Accepted for compatibility: walkers are batched on the GPU, not spread over processes.")
    group.add_argument("--mpi", dest="mpi", default=False, action="store_true", help="Accepted for compatibility (see --ncores).")
    parser.add_argument('--device', default=0, type=int, help='CUDA device ordinal of this process.')
    parser.add_argument('-v', '--verbose', action='store_true', help='Print convergence statistics and a progress bar.')
    parser.add_argument('--clever', action='store_true',
                        help='Tiered chunks from the most variable 800/200/50-day windows of the whole light curve; overrides '
                             '--quarters, --subsample and --daterange.')
    parser.add_argument('--npoints', type=int, default=600)
    parser.add_argument('--acf_prior', action='store_true')
    parser.add_argument('--sap', action='store_true')
    parser.add_argument('--filter', action='store_true')
    parser.add_argument('--bestchunk', nargs='+', type=int, default=None)
    parser.add_argument('--tag', default=None)
    parser.add_argument('--quarters', nargs='+', type=int, default=None)
    parser.add_argument('--sampler', choices=['mnest', 'emcee3', 'polychord'], default='emcee3', help='Sampler; only emcee3 is available here.')
    parser.add_argument('--resultsdir', default='results', help='Output directory for samples, light curve and priors.')
    parser.add_argument('-n', '--ndays', default=None, type=int, help='Use only the first NDAYS days of the light curve.')
    parser.add_argument('--subsample', default=30, type=int, help='Keep one point in SUBSAMPLE (random, seeded with the star id).')
    parser.add_argument('--chunksize', default=300, type=int, help='Largest number of points per independent chunk.')
    parser.add_argument('--nwalkers', default=500, type=int, help='Ensemble size.')
    parser.add_argument('--iter_chunksize', default=50, type=int, help='Steps per block between convergence checks.')
    parser.add_argument('--targetn', default=6, type=int, help='Chain length, in autocorrelation times, that counts as converged.')
    parser.add_argument('--nburn', default=2, type=int, help='Autocorrelation times discarded as burn-in.')
    parser.add_argument('--maxiter', default=50, type=int, help='Upper limit on the number of blocks.')
    parser.add_argument('-o', '--overwrite', action='store_true',
                        help='Start from scratch; without it a stored chain of the same star is resumed.')
    parser.add_argument('--daterange', nargs=2, type=float, help='Restrict the fit to times between the two values.')
    parser.add_argument('--mixedmoves', action='store_true', default=True,
                        help='KDE / differential-evolution / snooker moves in proportion 2:2:1 (default); otherwise KDE only.')
    parser.add_argument('--nlive', default=1000, type=int, help='Live points (MultiNest only).')
    parser.add_argument('--test', action='store_true', help="Dry run for the nested samplers; ignored by emcee3.")
    parser.add_argument('--offline', action='store_true', help='kplr offline mode.')
    parser.add_argument('--altmodel', action='store_true')
    parser.add_argument('--nochunks', action='store_true')
    parser.add_argument('--seed', default=None, type=int, help='Seed of the sampler (not in the reference).')
    return parser


def main(argv=None):
    args = vars(build_parser().parse_args(argv))
    args.pop('mpi')
    args.pop('n_cores')
    args['pool'] = None                      # fit_emcee3 -> EnsemblePool: one launch per proposal set
    stars = args.pop('stars')
    sampler = args.pop('sampler')
    N = len(stars)
    status = 0
    for i, ix in enumerate(stars):
        print('{} of {}: {}'.format(i + 1, N, ix))
        try:
            if sampler == 'emcee3':
                _fit_emcee3(ix, **args)
            else:
                raise NotImplementedError('--sampler {}: MultiNest / PolyChord are not installed here; the model exposes '
                                          'mnest_prior / mnest_loglike / polychord_* for them.'.format(sampler))
        except Exception:
            import traceback

            traceback.print_exc()
            logging.error('Error with {}; traceback above.'.format(ix))
            status = 1
    return status


if __name__ == '__main__':
    sys.exit(main())
