"""Drop-in module paths of the reference (`from src.models... import ...` in run_animate.py:13-16 / run_edit.py:13-16)
backed by the mimo_b200 engine. Overlay `src/models` and `src/pipelines` on a checkout of menyifang/MIMO (its own
`src/utils`, `tools/` and entry scripts stay as they are); see INTEGRATION.md."""
