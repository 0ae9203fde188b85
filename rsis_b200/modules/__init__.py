"""Drop-in counterparts of the reference's `src/modules/{vision,clstm,model}.py`."""
