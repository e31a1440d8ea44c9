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
    parser = argparse.ArgumentParser(description='Fit aigrain simulations or kepler light curves with GP rotation model.')
    parser.add_argument('stars', nargs='+', type=int, help='Stars to fit.')
    datagroup = parser.add_mutually_exclusive_group()
    datagroup.add_argument("--aigrain", dest="aigrain", action='store_true', help="Use Aigrain simulations.")
    datagroup.add_argument("--kepler", dest="kepler", action='store_true', help="Use Kepler data.")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--ncores", dest="n_cores", default=1, type=int,
                       help="Accepted for compatibility: walkers are batched on the GPU, not spread over processes.")
    group.add_argument("--mpi", dest="mpi", default=False, action="store_true", help="Accepted for compatibility (see --ncores).")
    parser.add_argument('--device', default=0, type=int, help='CUDA device ordinal of this process.')
    parser.add_argument('-v', '--verbose', action='store_true', help='Have the sampler talk to you.')
    parser.add_argument('--clever', action='store_true',
                        help='Subsample in a tiered way but use whole light curve. Supersedes "quarters", "subsample", '
                             'and "daterange" arguments.')
    parser.add_argument('--npoints', type=int, default=600)
    parser.add_argument('--acf_prior', action='store_true')
    parser.add_argument('--sap', action='store_true')
    parser.add_argument('--filter', action='store_true')
    parser.add_argument('--bestchunk', nargs='+', type=int, default=None)
    parser.add_argument('--tag', default=None)
    parser.add_argument('--quarters', nargs='+', type=int, default=None)
    parser.add_argument('--sampler', choices=['mnest', 'emcee3', 'polychord'], default='emcee3', help='Which sampling method to use.')
    parser.add_argument('--resultsdir', default='results', help='Directory in which to store results.')
    parser.add_argument('-n', '--ndays', default=None, type=int, help='Number of days (from beginning) of light curve to use for fitting.')
    parser.add_argument('--subsample', default=30, type=int, help='subsampling factor.')
    parser.add_argument('--chunksize', default=300, type=int, help='Size (in number of points) of the chunks into which to split light curve.')
    parser.add_argument('--nwalkers', default=500, type=int, help='Number of walkers (for emcee3)')
    parser.add_argument('--iter_chunksize', default=50, type=int, help='Number of emcee3 steps after which to check for convergence.')
    parser.add_argument('--targetn', default=6, type=int, help='Number of autocorrelation times after which to say fit has converged.')
    parser.add_argument('--nburn', default=2, type=int, help='Number of autocorrelation times to toss out as burn-in.')
    parser.add_argument('--maxiter', default=50, type=int, help='Maximum number of times to repeat <iter_chunksize> steps of emcee3.')
    parser.add_argument('-o', '--overwrite', action='store_true',
                        help='Overwrite existing fit.  If this is not set and a fit has been done/started before, it will '
                             'restart from previous stopping point.')
    parser.add_argument('--daterange', nargs=2, type=float, help='Beginning and end of specified date range of fit.')
    parser.add_argument('--mixedmoves', action='store_true', default=True,
                        help='If this is set, then a 2:2:1 mix of KDEMove, DEMove and DESnookerMove are used in emcee3 fit. '
                             'Otherwise, just KDEMove is used.')
    parser.add_argument('--nlive', default=1000, type=int, help='Number of live points (for multinest)')
    parser.add_argument('--test', action='store_true', help="Test, don't actually run. Flag has no effect if sampler is set to emcee3.")
    parser.add_argument('--offline', action='store_true', help='enable offline kplr data access')
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
