"""Type aliases used by the reference (conv.py:16-22, block.py:22-24, mapper.py:21-23)."""
from typing import Optional, Tuple, Union

from torch import Tensor

Adj = Union[Tensor, object]
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
PairOptTensor = Tuple[Optional[Tensor], Optional[Tensor]]
Size = Optional[Tuple[int, int]]
NoneType = Optional[Tensor]
