"""TEST INFRASTRUCTURE ONLY -- the reference build recipe, its ctypes bindings and the CPU restatement of the hot path.
Nothing under rebvo_b200/ or include/ imports, links or executes anything from this package."""
