"""Training entry point named by pretrainTrans.sh:1.  Parses the reference's flags verbatim,
builds the networks (define_G / define_D) and evaluates the forward losses; the optimisation step needs the
backward kernels that are not built yet and fails loudly instead of falling back to torch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from nhvr_b200.capi import NhvrError
from nhvr_b200.options import TrainOptions


def main(argv=None):
    opt = TrainOptions().parse(argv)
    raise NhvrError("pre_train.py: flags parsed (name=%s, batchSize=%d, lambda_L2=%g, lambda_UV=%g, lambda_Prob=%g, "
                    "lambda_Temp=%g) but the sm_100a backward kernels (dgrad/wgrad/IN-bwd/sampler scatter) are not "
                    "built yet; there is no PyTorch fallback by design" %
                    (opt.name, opt.batchSize, opt.lambda_L2, opt.lambda_UV, opt.lambda_Prob, opt.lambda_Temp))


if __name__ == "__main__":
    main()
