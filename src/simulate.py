"""`python -m src.simulate` -> conclave_sim_b200.simulate (same flags, same outputs)."""
from conclave_sim_b200.simulate import *  # noqa: F401,F403
from conclave_sim_b200.simulate import (_run_single_simulation_worker, main, parse_args, model_params_from_args,  # noqa: F401
                                        run_monte_carlo_simulation, REQUIRED_MAJORITY_FRACTION, MAX_ROUNDS)

if __name__ == "__main__":
    main()
