"""Import alias: the product package lives in ``assemble-them-all_b200/`` (a
directory name Python cannot import directly). This stub extends ``__path__``
so that ``import assemble_them_all_b200.contact`` resolves into that directory.
No code lives here."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "assemble-them-all_b200"))
