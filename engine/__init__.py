"""Drop-in import paths of the reference (`engine.engine`, `engine.mcts`, `engine.games.<game>.<backend>`,
`engine.value_functions`, `engine.policy_functions`): aliases of the zeroclone_b200 package."""
