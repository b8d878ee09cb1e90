from engine._alias import alias

alias(__name__, "zeroclone_b200.games.connect4.c4_backend")
