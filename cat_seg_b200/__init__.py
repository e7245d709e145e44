"""Importable alias of the ``cat-seg_b200/`` package directory.

The product lives in ``cat-seg_b200/`` (a hyphen is not a legal Python identifier), so this thin
package points its ``__path__`` there: ``import cat_seg_b200.aggregator`` loads
``cat-seg_b200/aggregator.py``.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cat-seg_b200"))
