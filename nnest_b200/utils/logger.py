"""Logger and run-directory layout.  Same on-disk contract as the reference's nnest/utils/logger.py
(create_logger :9-22, get_or_create_run_dir :38-75): <log_dir>/runN/{info,results,chains,checkpoint,plots}."""
import logging
import os
import sys

SUBDIRS = ('info', 'results', 'chains', 'checkpoint', 'plots')


def create_logger(module_name, level=logging.INFO):
    logger = logging.getLogger(module_name)
    for hd in list(logger.handlers):
        logger.removeHandler(hd)
    logger.setLevel(level)
    hd = logging.StreamHandler(sys.stdout)
    hd.setLevel(level)
    hd.setFormatter(logging.Formatter('[%s] [%%(levelname)s] %%(message)s' % module_name))
    logger.addHandler(hd)
    return logger


def get_or_create_run_dir(run_dir, append_run_num=True):
    """Returns the dict of sub-directories plus 'created'.  An existing <run_dir>/info means "resume here";
    otherwise a new run<k> (k = number of existing sub-directories + 1) is made when append_run_num."""
    created = not os.path.isdir(os.path.join(run_dir, 'info'))
    if created:
        os.makedirs(run_dir, exist_ok=True)
        if append_run_num:
            k = sum(os.path.isdir(os.path.join(run_dir, e)) for e in os.listdir(run_dir)) + 1
            run_dir = os.path.join(run_dir, 'run%s' % k)
            if not os.path.isdir(run_dir):
                print('Creating directory for new run %s' % run_dir)
        if not os.path.isdir(os.path.join(run_dir, 'info')):
            for sub in SUBDIRS:
                os.makedirs(os.path.join(run_dir, sub))
    else:
        print('Using old directory %s' % run_dir)
    out = {sub: os.path.join(run_dir, sub) for sub in SUBDIRS}
    out['run_dir'] = run_dir
    out['created'] = created
    return out
