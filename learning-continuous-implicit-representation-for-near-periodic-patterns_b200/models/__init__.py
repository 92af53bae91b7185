"""Drop-in replacement of the reference's ``models`` package for the NPP-Net training hot path.

Put this directory's parent (``learning-..._b200/``) on ``PYTHONPATH`` *before* the reference root: the reference
scripts append their own root to the END of ``sys.path`` (NPP_completion/train.py:4), so this package shadows
``models`` while ``loaders``, ``options``, ``utils`` and ``externel_lib`` still resolve to the reference.
See INTEGRATION.md.
"""
