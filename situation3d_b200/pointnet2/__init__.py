"""Drop-in for the reference package ``lib/pointnet2`` (module names and public
symbols identical: ``_ext``, ``pointnet2_utils``, ``pointnet2_modules``,
``pytorch_utils``)."""
