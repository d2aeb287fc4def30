from .gptq import *  # noqa: F401,F403
from .gptq import _accumulate_hessian, _gptq, _gptq_quantize  # noqa: F401
from .hqq import *  # noqa: F401,F403
from .hqq import _hqq_quantize  # noqa: F401
from .rtn import *  # noqa: F401,F403
from .rtn import _quantize_bias, _rtn_quantize  # noqa: F401
