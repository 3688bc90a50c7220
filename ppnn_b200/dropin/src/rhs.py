from ppnn_b200.rhs import (burgers2DFu, burgers2DFv, burgers2Dpu, burgers2Dpv,  # noqa: F401
                           burgers2Ddiffuseu, burgers2Ddiffusev, rdFu, rdFv, rdpu, rdpv,
                           diffu, diffv, convecu, convecv, ks1d, burgers2DWu, burgers2DWv)
