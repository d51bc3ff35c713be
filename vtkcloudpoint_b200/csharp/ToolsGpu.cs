// ToolsGpu.cs -- drop-ins for the statistics / matching helpers that sit on either side of DBSCAN and ICP in MainForm and Tools
// (SURVEY.md 8f).  Same inputs and post-state as the C# they replace; the arithmetic runs in libvpc.  Source only (no .NET here).
using System;
using System.Collections.Generic;

namespace vtkPointCloud
{
    static class ToolsGpu
    {
        // Tools.GetClusList (Tools.cs:162-195) + the two Tools.getCircles calls of CompleteWork3 (FrmMain.cs:1521-1540).
        // rawData carries clusterId 0..clusterSum; centers / centers2D / circles / circles2D are filled exactly as the C# fills them.
        public static void ClusterStatistics(IntPtr ctx, List<Point3D> rawData, int clusterSum, List<Point3D> centers, List<Point3D> centers2D,
                                             List<Point2D> circles, List<Point2D> circles2D)
        {
            int n = rawData.Count, k1 = clusterSum + 1;
            int[] cid = new int[n];
            double[] xyz = new double[3 * n], mx = new double[n], my = new double[n];
            for (int i = 0; i < n; i++)
            {
                Point3D p = rawData[i];
                cid[i] = p.clusterId; xyz[i] = p.X; xyz[n + i] = p.Y; xyz[2 * n + i] = p.Z; mx[i] = p.motor_x; my[i] = p.motor_y;
            }
            double[] means = new double[5 * k1], c3 = new double[3 * k1], c2 = new double[3 * k1];
            int[] counts = new int[k1], s3 = new int[k1], s2 = new int[k1];
            NativeMethods.Check(ctx, NativeMethods.vpc_cluster_stats(ctx, cid, n, clusterSum, xyz, mx, my, means, counts, c3, s3, c2, s2));
            for (int c = 1; c <= clusterSum; c++)
            {
                if (counts[c] == 0) continue;                                                   // Tools.cs:191
                centers.Add(new Point3D(means[c], means[k1 + c], means[2 * k1 + c], c, true));  // :192
                centers2D.Add(new Point3D(means[3 * k1 + c], means[4 * k1 + c], 0, c, true));   // :193
            }
            for (int c = 1; c <= clusterSum; c++)                                               // Tools.cs:398-407
            {
                if (s3[c] == 1) { Point2D q = new Point2D(c3[c], c3[k1 + c]); q.radius = c3[2 * k1 + c]; q.clusID = c; circles.Add(q); }
                else if (s3[c] < 0) throw new MException("getCircles: cluster " + c + " cannot be processed (status " + s3[c] + ")");
                if (s2[c] == 1) { Point2D q = new Point2D(c2[c], c2[k1 + c]); q.radius = c2[2 * k1 + c]; q.clusID = c; circles2D.Add(q); }
            }
        }

        // MainForm.refreshClusList's per-point query (FrmMain.cs:3446-3467): returns the clusterId of the nearest truth within the radius, 0 if none
        public static int[] NearestTruth(IntPtr ctx, List<Point3D> trues, List<Point3D> rawData, double clusterRadius)
        {
            int m = trues.Count, n = rawData.Count;
            double[] tx = new double[m], ty = new double[m], px = new double[n], py = new double[n];
            int[] tid = new int[m], id = new int[n];
            for (int s = 0; s < m; s++) { tx[s] = trues[s].tmp_X; ty[s] = trues[s].tmp_Y; tid[s] = trues[s].clusterId; }
            for (int i = 0; i < n; i++) { px[i] = rawData[i].motor_x; py[i] = rawData[i].motor_y; }
            NativeMethods.Check(ctx, NativeMethods.vpc_nearest_truth_2d(ctx, tx, ty, tid, m, px, py, n, clusterRadius, id));
            return id;
        }
    }
}
