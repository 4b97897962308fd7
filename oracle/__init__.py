"""CPU oracle for tgp's Reduce + Connect hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``torch-geometric-pool_b200/tgp_b200``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the CPU baseline.

Pinning status: the restated PyG primitives (``pyg_shim``) and the restated
reference path (``ref_path``) are pinned two ways:

* against every hand value the reference's own tests hold for this path
  (re-typed in ``tests/test_oracle_pins.py``);
* against outputs of the reference's *own* source files executed in this
  container through ``oracle/ref_import.py`` (stub ``torch_geometric`` /
  ``torch_scatter`` modules exposing only the restated primitives), committed as
  fixtures under ``tests/golden/`` with the generating script.

The PyG primitives themselves (torch_geometric >=2.6,<3, torch_scatter 2.1.2) are
NOT present in this image, so their semantics are restated from the published
behaviour of that pinned range; coarse-edge ordering / relabel numbering are
pinned only by those restated rules -> "parity partially
unpinned" for exactly those rules, stated again in DESIGN.md.
"""
