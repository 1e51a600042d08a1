"""Drop-in for the reference's src/modeler_spark_driver.py: ``python -m
time_series_spark_b200.modeler_driver <config.yaml>`` (no spark-submit, no SparkSession)."""
import sys

import yaml

from .jobs.prophet_modeler import ProphetModeler

if __name__ == "__main__":
    if len(sys.argv) != 2:
        print("arg1 must be the config YAML")
        sys.exit(1)
    with open(sys.argv[1]) as file:
        config = yaml.safe_load(file)
    print(f"config: {config}")
    ProphetModeler.model(None, config)
