"""Model-size table (MB of parameters = bytes moved per PS/worker exchange) used by the network-cost model.
The values are the checkpoint sizes listed in the reference's model/model_factory.py:19-55 (data, not
code); the reference never imports that module — here it feeds rlgs_netcost_inputs.model_mb."""

model_sizes = {
    '4_layers_brnn': 1300, 'transformer': 1100, '1_layer_bilstm_opennmt': 900, 'BERT_Chinese': 350,
    '2_layers_lstm_gigaword': 330, 'mobilenet_v1_025': 15, 'googlenet': 26, 'inception2': 43, 'inception3': 104,
    'inception4': 176, 'alexnet': 233, 'vgg11': 519, 'vgg19': 549, 'vgg16': 528, 'resnet50': 97,
    'resnet101': 555, 'resnet152': 737,
}

cnn_models = ['mobilenet_v1_025', 'googlenet', 'inception2', 'inception3', 'inception4', 'alexnet', 'vgg11', 'vgg16',
              'vgg19', 'resnet50', 'resnet101', 'resnet152']


def size_mb(name):
    """MB for a model name, 0.0 when unknown (e.g. the Philly traces' `model` column holds a GPU type)."""
    return float(model_sizes.get(str(name), 0.0))
