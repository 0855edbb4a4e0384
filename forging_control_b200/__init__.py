"""forging_control_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path of
marcowus/forging-control: the fused MPC-loss roll-out + reverse sweep and the closed-loop RK4
deployment roll-out, behind the reference's own Python API (see ``Functions.py``)."""
from . import _native
from .Functions import (Data, FNNModel, FeasibilityRecovery, LSTMModel, MPCLoss, NeuralNetwork,
                        mpc_loss_native, pack_weights, lstm_shadow_native)
from .dataset import DeviceSequenceLoader, build_windows
from .closed_loop import closed_loop_device, closed_loop_rollout, tvp_reference_table
from .surrogate import DeviceAdamW, SurrogateNeuralNetwork, lstm_window
from .distributed import (FlatGradBucket, allreduce_loss_and_grads, shard_bounds, sharded_surrogate_step,
                          sharded_training_step)

__all__ = ["FNNModel", "LSTMModel", "MPCLoss", "NeuralNetwork", "FeasibilityRecovery", "Data",
           "mpc_loss_native", "pack_weights", "closed_loop_device", "closed_loop_rollout",
           "tvp_reference_table", "shard_bounds", "allreduce_loss_and_grads", "sharded_training_step",
           "lstm_shadow_native", "DeviceSequenceLoader", "build_windows", "install",
           "sharded_surrogate_step", "FlatGradBucket", "DeviceAdamW", "SurrogateNeuralNetwork", "lstm_window", "install_surrogate"]


def install(reference_functions_module) -> None:
    """Swap the hot-path classes into an already imported reference ``Functions`` module so that the
    reference ``Main.py`` (``from Functions import ...``, Main.py:19) picks up the CUDA path:

        import Functions, forging_control_b200 as fb; fb.install(Functions)
    """
    ref = reference_functions_module
    ref.FNNModel = FNNModel
    ref.LSTMModel = LSTMModel
    ref.MPCLoss = MPCLoss
    ref.NeuralNetwork.train_model = staticmethod(NeuralNetwork.train_model)
    ref.NeuralNetwork.validate_model = staticmethod(NeuralNetwork.validate_model)
    ref.NeuralNetwork.loop = staticmethod(NeuralNetwork.loop)
    ref.FeasibilityRecovery.NN_make_step = staticmethod(FeasibilityRecovery.NN_make_step)


def install_surrogate(reference_model_nn_functions_module) -> None:
    """The same for the surrogate-training script (``Model_NN/Main.py:26`` imports ``Model_NN/Functions.py``):

        import Functions, forging_control_b200 as fb; fb.install_surrogate(Functions)
    """
    ref = reference_model_nn_functions_module
    ref.LSTMModel = LSTMModel
    ref.NeuralNetwork.train_model = staticmethod(SurrogateNeuralNetwork.train_model)
    ref.NeuralNetwork.validate_model = staticmethod(SurrogateNeuralNetwork.validate_model)
