from engine._alias import alias

alias(__name__, "zeroclone_b200.engine")
