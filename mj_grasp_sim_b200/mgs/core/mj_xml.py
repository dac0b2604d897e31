"""MjXml protocol (/root/reference/mgs/core/mj_xml.py:21-30): entities emit (xml fragment, assets)."""
from typing import Any, Dict, Protocol, Tuple


class MjXml(Protocol):
    def to_xml(self) -> Tuple[str, Dict[str, Any]]:
        ...
